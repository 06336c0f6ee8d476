#!/bin/bash
# round-2 first hardware pass of the split-fp16 family: unit parity, network goldens, step tables
set +e
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2a_smi.txt 2>&1
T="timeout 600 python -m pytest -q -p no:cacheprovider"
$T tests/test_gpu_split16.py -k "elementwise or bn_bwd" > gpurun_out/r2a_elem.log 2>&1; echo "elem rc=$?"
$T tests/test_gpu_split16.py -k "fprop" > gpurun_out/r2a_fprop.log 2>&1; echo "fprop rc=$?"
$T tests/test_gpu_split16.py -k "dgrad" > gpurun_out/r2a_dgrad.log 2>&1; echo "dgrad rc=$?"
$T tests/test_gpu_split16.py -k "wgrad" > gpurun_out/r2a_wgrad.log 2>&1; echo "wgrad rc=$?"
$T tests/test_gpu_parity.py -k "network_vs_reference_golden and f16x3" > gpurun_out/r2a_net.log 2>&1; echo "net rc=$?"
timeout 600 python tools/step_table.py 32 f16x3 > gpurun_out/r2a_step_f16x3.md 2> gpurun_out/r2a_step_f16x3.err; echo "step16 rc=$?"
timeout 600 python tools/step_table.py 32 tf32x3 > gpurun_out/r2a_step_tf32x3.md 2> gpurun_out/r2a_step_tf32x3.err; echo "step32 rc=$?"
tail -5 gpurun_out/r2a_*.log
head -30 gpurun_out/r2a_step_f16x3.md
