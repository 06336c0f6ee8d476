#!/bin/bash
set +e
mkdir -p gpurun_out
for L in "128 16 16 256 1024 1 1" "128 32 32 128 512 1 1" "128 16 16 1024 256 1 1" "128 64 64 64 256 1 1" "128 16 16 256 256 3 1" "128 64 64 256 64 1 1" "128 64 64 64 64 3 1"; do
  python tools/one_conv16.py $L fprop 7 2>&1 | tail -1
done
python -m pytest tests -q -x -m gpu > gpurun_out/r2r_gpu_tests.log 2>&1; echo "gpu tests rc=$?"
tail -4 gpurun_out/r2r_gpu_tests.log
python bench.py > gpurun_out/r2r_bench.json 2> gpurun_out/r2r_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2r_bench.json'))
print(d['ms_per_step'], d['value'], d['e2e'], d['roofline']['achieved'], d['roofline']['conv_family'], d['gpu_launches'])
PY
