#!/bin/bash
set +e
mkdir -p gpurun_out
python tools/dbg_refiner.py 2>&1 | tail -8
timeout 2400 python -m pytest tests/ -q -m gpu -p no:cacheprovider -s --deselect tests/test_gpu_parity.py::test_refiner_train_loop_and_checkpoint_gpu > gpurun_out/r2h_gpu_tests.log 2>&1; echo "gpu tests rc=$?"
tail -n 8 gpurun_out/r2h_gpu_tests.log
