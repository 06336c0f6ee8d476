#!/bin/bash
set +e
mkdir -p gpurun_out
python -m pytest tests -q -x -m gpu > gpurun_out/r2t_gpu_tests.log 2>&1; echo "gpu tests rc=$?"
tail -4 gpurun_out/r2t_gpu_tests.log
python tools/bench_aux.py 2>/dev/null | grep patch_sample | cut -c1-230
python bench.py > gpurun_out/r2t_bench.json 2> gpurun_out/r2t_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2t_bench.json'))
print(d['ms_per_step'], d['value'], d['e2e'], d['roofline']['achieved'], d['roofline']['conv_family']['achieved'], d['gpu_launches'])
PY
