#!/bin/bash
# round-2 second hardware pass: epilogue v2 (TMA store), dynamic activation scales, size parity
set +e
mkdir -p gpurun_out
T="timeout 900 python -m pytest -q -p no:cacheprovider"
$T tests/test_gpu_split16.py > gpurun_out/r2c_split16.log 2>&1; echo "split16 rc=$?"
$T tests/test_gpu_parity.py -k "network_vs_reference_golden" > gpurun_out/r2c_net.log 2>&1; echo "net rc=$?"
$T tests/test_gpu_sizes.py -s > gpurun_out/r2c_sizes.log 2>&1; echo "sizes rc=$?"
for cfg in "128 64 64 64 256 1 1 fprop" "128 16 16 256 256 3 1 fprop" "128 64 64 256 64 1 1 dgrad" "128 32 32 128 512 1 1 fprop" "128 64 64 64 64 3 1 fprop"; do
  timeout 120 python tools/one_conv16.py $cfg 2>&1 | tail -1
done
timeout 600 python tools/step_table.py 32 f16x3 > gpurun_out/r2c_step_f16x3.md 2> gpurun_out/r2c_step_f16x3.err; echo "step16 rc=$?"
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err; echo "bench rc=$?"
tail -n 4 gpurun_out/r2c_split16.log gpurun_out/r2c_net.log gpurun_out/r2c_sizes.log
head -n 32 gpurun_out/r2c_step_f16x3.md
cat gpurun_out/r2c_bench.json | head -c 1500
