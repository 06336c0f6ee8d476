#!/bin/bash
set +e
mkdir -p gpurun_out
python -m pytest tests/test_gpu_split16.py -q -x -k "fprop_vs_emulation or bench_layer" > gpurun_out/r3c_unit.log 2>&1; echo "unit rc=$?"; tail -2 gpurun_out/r3c_unit.log
for L in "128 64 64 64 256 1 1" "128 16 16 256 1024 1 1" "128 32 32 128 512 1 1"; do
  echo "== $L"; EPB_C16_PROBE=32 python tools/one_conv16.py $L fprop 5 2>&1 | tail -7 | head -4
done
