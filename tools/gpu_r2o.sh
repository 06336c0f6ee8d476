#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_split16.py -q -x -k "bn_bwd or softargmax_bwd" > gpurun_out/r2o_tests.log 2>&1; echo "tests rc=$?"
python tools/timeline.py 32 gpurun_out/r2o_timeline.csv > gpurun_out/r2o_timeline.md 2> gpurun_out/r2o_timeline.err; echo "timeline rc=$?"
python bench.py > gpurun_out/r2o_bench.json 2> gpurun_out/r2o_bench.err; echo "bench rc=$?"
tail -2 gpurun_out/r2o_tests.log; cat gpurun_out/r2o_timeline.md; tail -3 gpurun_out/r2o_timeline.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2o_bench.json'))
print(d['ms_per_step'], d['value'], d['e2e'], d['roofline']['achieved'], d['roofline']['conv_family']['achieved'], d['gpu_launches'])
PY
