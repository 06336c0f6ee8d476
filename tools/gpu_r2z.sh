#!/bin/bash
set +e
mkdir -p gpurun_out
python -m pytest tests/test_gpu_split16.py -q -x -k "conv16 or fprop or dgrad" > gpurun_out/r2z_unit.log 2>&1; echo "unit rc=$?"; tail -2 gpurun_out/r2z_unit.log
for L in "128 64 64 64 256 1 1" "128 16 16 256 1024 1 1" "128 32 32 128 512 1 1" "128 16 16 1024 256 1 1" "128 16 16 256 256 3 1" "128 64 64 64 64 3 1" "128 64 64 256 64 1 1"; do
  echo "== $L"; EPB_C16_PROBE=32 python tools/one_conv16.py $L fprop 5 2>&1 | tail -7 | head -4
done
python bench.py --no-cpu-baseline > gpurun_out/r2z_bench.json 2> gpurun_out/r2z_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2z_bench.json'))
print(d['ms_per_step'], d['value'], d['e2e'], d['roofline']['achieved'], d['roofline']['conv_family'])
PY
