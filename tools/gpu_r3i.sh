#!/bin/bash
# last evidence pass of the round: ncu launch list of one step and one full capture of the dominant kernel, final code
set +e
mkdir -p gpurun_out
timeout 400 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r3i_launches.csv python tools/profile_step.py > gpurun_out/r3i_profile_step.log 2>&1
python tools/summarize_launches.py gpurun_out/r3i_launches.csv > gpurun_out/r3i_launches.md 2>&1; head -14 gpurun_out/r3i_launches.md
timeout 200 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"conv16_kernel" -s 2 -c 1 -o gpurun_out/r3i_c16 -f python tools/one_conv16.py 128 16 16 256 256 3 1 fprop > gpurun_out/r3i_ncu_c16.log 2>&1
python tools/ncu_summary.py gpurun_out/r3i_c16.ncu-rep 14 > gpurun_out/r3i_sum_c16.txt 2>&1; head -22 gpurun_out/r3i_sum_c16.txt
