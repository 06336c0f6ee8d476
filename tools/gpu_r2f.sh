#!/bin/bash
set +e
mkdir -p gpurun_out
python tools/one_bn16.py 524288 256 2>&1 | tail -3
python tools/one_bn16.py 524288 64 2>&1 | tail -3
python tools/one_bn16.py 32768 1024 2>&1 | tail -3
timeout 300 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"bn_bwd_reduce_mx" -s 2 -c 1 -o gpurun_out/r2f_red -f python tools/one_bn16.py 524288 256 2 > gpurun_out/r2f_ncu_red.log 2>&1
python tools/ncu_summary.py gpurun_out/r2f_red.ncu-rep 16 > gpurun_out/r2f_sum_red.txt 2>&1
timeout 300 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"conv16_kernel" -s 2 -c 1 -o gpurun_out/r2f_c16 -f python tools/one_conv16.py 128 64 64 64 256 1 1 fprop > gpurun_out/r2f_ncu_c16.log 2>&1
python tools/ncu_summary.py gpurun_out/r2f_c16.ncu-rep 24 > gpurun_out/r2f_sum_c16.txt 2>&1
T="timeout 1200 python -m pytest -q -p no:cacheprovider"
$T tests/test_gpu_split16.py -k "bench_layer" > gpurun_out/r2f_layers.log 2>&1; echo "layers rc=$?"
$T tests/test_gpu_sizes.py -s > gpurun_out/r2f_sizes.log 2>&1; echo "sizes rc=$?"
$T tests/test_gpu_parity.py -s -k "final_preds or eight_point or polynomial or occluder or refiner or input_pipeline" > gpurun_out/r2f_new.log 2>&1; echo "new rc=$?"
tail -n 4 gpurun_out/r2f_layers.log gpurun_out/r2f_new.log
grep -E "refiner train|heat-maps|gradients|passed|failed|^E  " gpurun_out/r2f_sizes.log gpurun_out/r2f_new.log | head -30
cat gpurun_out/r2f_sum_red.txt | head -45
