#!/bin/bash
set +e
for L in "128 64 64 64 256 1 1" "128 16 16 256 1024 1 1"; do
  for W in fprop dgrad; do
    echo "== $L $W"; EPB_C16_PROBE=32 python tools/one_conv16.py $L $W 5 2>&1 | tail -8
  done
done
