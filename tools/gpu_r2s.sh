#!/bin/bash
set +e
mkdir -p gpurun_out
for L in "128 16 16 256 1024 1 1" "128 32 32 128 512 1 1" "128 16 16 1024 256 1 1" "128 64 64 64 256 1 1" "128 64 64 256 64 1 1" "128 64 64 256 1024 1 1"; do
  for P in 0 16; do
    for W in fprop dgrad; do
      echo -n "probe $P: "; EPB_C16_PROBE=$P python tools/one_conv16.py $L $W 7 2>&1 | tail -1
    done
  done
done
EPB_C16_PROBE=16 python -m pytest tests/test_gpu_split16.py -q -x -k "conv16" > gpurun_out/r2s_tests.log 2>&1; echo "tests(lsu) rc=$?"; tail -2 gpurun_out/r2s_tests.log
