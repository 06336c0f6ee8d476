#!/bin/bash
# final validation of the round: full GPU suite, smoke(), both bench arms
set +e
mkdir -p gpurun_out
python -m pytest tests -q -x -m gpu > gpurun_out/r3b_gpu_tests.log 2>&1; echo "gpu tests rc=$?"; tail -3 gpurun_out/r3b_gpu_tests.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r3b_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r3b_smoke.log
timeout 900 python bench.py --impl reference > gpurun_out/r3b_bench_ref.json 2> gpurun_out/r3b_bench_ref.err; echo "ref rc=$?"
timeout 900 python bench.py > gpurun_out/r3b_bench.json 2> gpurun_out/r3b_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r3b_bench.json')); r=json.load(open('gpurun_out/r3b_bench_ref.json'))
print(d['ms_per_step'], d['value'], d['e2e']['value'], d['roofline']['achieved'], d['roofline']['executed_frac'], d['roofline']['conv_family']['achieved'], 'ref', r['value'], r['cpu_baseline']['cores'])
PY
