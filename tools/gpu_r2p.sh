#!/bin/bash
set +e
mkdir -p gpurun_out
python tools/one_conv16.py 128 16 16 256 1024 1 1 fprop 2>&1 | tail -2
python tools/one_conv16.py 128 32 32 128 512 1 1 fprop 2>&1 | tail -2
python tools/one_conv16.py 128 16 16 1024 256 1 1 fprop 2>&1 | tail -2
timeout 300 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"conv16_kernel" -s 2 -c 1 -o gpurun_out/r2p_a -f python tools/one_conv16.py 128 16 16 256 1024 1 1 fprop > gpurun_out/r2p_ncu_a.log 2>&1
python tools/ncu_summary.py gpurun_out/r2p_a.ncu-rep 30 > gpurun_out/r2p_sum_a.txt 2>&1
timeout 300 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"conv16_kernel" -s 2 -c 1 -o gpurun_out/r2p_b -f python tools/one_conv16.py 128 32 32 128 512 1 1 fprop > gpurun_out/r2p_ncu_b.log 2>&1
python tools/ncu_summary.py gpurun_out/r2p_b.ncu-rep 30 > gpurun_out/r2p_sum_b.txt 2>&1
cat gpurun_out/r2p_sum_a.txt | head -60
cat gpurun_out/r2p_sum_b.txt | head -40
