#!/bin/bash
set +e
for L in "128 64 64 64 256 1 1" "128 16 16 256 1024 1 1" "128 32 32 128 512 1 1" "128 16 16 256 256 3 1"; do
  for P in 32 44; do
    echo "== $L probe $P"; EPB_C16_PROBE=$P python tools/one_conv16.py $L fprop 5 2>&1 | tail -7
  done
done
