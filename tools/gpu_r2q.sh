#!/bin/bash
set +e
mkdir -p gpurun_out
for L in "128 16 16 256 1024 1 1" "128 32 32 128 512 1 1" "128 16 16 1024 256 1 1" "128 64 64 64 256 1 1" "128 16 16 256 256 3 1"; do
  for P in 0 8 12 1 9 2 10; do
    echo -n "probe $P: "; EPB_C16_PROBE=$P python tools/one_conv16.py $L fprop 7 2>&1 | tail -1
  done
done
