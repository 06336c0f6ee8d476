#!/bin/bash
# ncu captures of the split-fp16 kernels on three bench layers
set +e
mkdir -p gpurun_out
i=0
for cfg in "128 64 64 64 256 1 1 fprop" "128 16 16 256 256 3 1 fprop" "128 64 64 64 256 1 1 wgrad" "128 32 32 128 512 1 1 fprop"; do
  i=$((i+1))
  timeout 120 python tools/one_conv16.py $cfg 2>&1 | tail -1
  timeout 300 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"(conv16|wgrad16)_kernel" -s 2 -c 1 -o gpurun_out/r2b_one$i -f python tools/one_conv16.py $cfg > gpurun_out/r2b_ncu$i.log 2>&1
  python tools/ncu_summary.py gpurun_out/r2b_one$i.ncu-rep 24 > gpurun_out/r2b_sum$i.txt 2>&1
done
