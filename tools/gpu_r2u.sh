#!/bin/bash
set +e
mkdir -p gpurun_out
python -m pytest tests/test_gpu_split16.py -q -x -k "bn_bwd or elementwise" > gpurun_out/r2u_unit.log 2>&1; echo "unit rc=$?"; tail -3 gpurun_out/r2u_unit.log
python tools/one_bn16.py 524288 256 7 2>&1 | head -3
python tools/one_bn16.py 32768 1024 7 2>&1 | head -3
python -m pytest tests -q -x -m gpu > gpurun_out/r2u_gpu_tests.log 2>&1; echo "gpu tests rc=$?"; tail -3 gpurun_out/r2u_gpu_tests.log
python tools/step_table.py > gpurun_out/r2u_step_f16x3.md 2> gpurun_out/r2u_step.err; echo "step rc=$?"
sed -n 1,12p gpurun_out/r2u_step_f16x3.md
python bench.py > gpurun_out/r2u_bench.json 2> gpurun_out/r2u_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2u_bench.json'))
print(d['ms_per_step'], d['value'], d['e2e'], d['roofline']['achieved'], d['roofline']['conv_family']['achieved'], d['gpu_launches'])
PY
