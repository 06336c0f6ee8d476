#!/bin/bash
set +e
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -q -x -k "heatmap" > gpurun_out/r3f_unit.log 2>&1; echo "unit rc=$?"; tail -2 gpurun_out/r3f_unit.log
/usr/local/cuda/bin/compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py -q -x -p no:cacheprovider -k "heatmap_joint_loss_vs_oracle" 2>&1 | grep -E "ERROR SUMMARY|passed|failed" | tail -2
python tools/bench_aux.py 2>/dev/null | grep heatmap | cut -c1-200
