#!/bin/bash
set +e
mkdir -p gpurun_out
timeout 200 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"bn_bwd_partial_kernel" -s 6 -c 1 -o gpurun_out/r3j_part -f python tools/one_bn16.py 524288 256 2 > gpurun_out/r3j_ncu_part.log 2>&1
python tools/ncu_summary.py gpurun_out/r3j_part.ncu-rep 8 > gpurun_out/r3j_sum_part.txt 2>&1; head -20 gpurun_out/r3j_sum_part.txt
timeout 200 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"bn_bwd_apply_split_kernel" -s 6 -c 1 -o gpurun_out/r3j_apply -f python tools/one_bn16.py 524288 256 2 > gpurun_out/r3j_ncu_apply.log 2>&1
python tools/ncu_summary.py gpurun_out/r3j_apply.ncu-rep 6 > gpurun_out/r3j_sum_apply.txt 2>&1; head -16 gpurun_out/r3j_sum_apply.txt
