#!/bin/bash
set +e
mkdir -p gpurun_out
T="timeout 1200 python -m pytest -q -p no:cacheprovider"
$T tests/test_gpu_split16.py > gpurun_out/r2d_split16.log 2>&1; echo "split16 rc=$?"
$T tests/test_gpu_sizes.py -s > gpurun_out/r2d_sizes.log 2>&1; echo "sizes rc=$?"
$T tests/test_gpu_parity.py > gpurun_out/r2d_parity.log 2>&1; echo "parity rc=$?"
timeout 600 python tools/step_table.py 32 f16x3 > gpurun_out/r2d_step_f16x3.md 2> gpurun_out/r2d_step_f16x3.err; echo "step16 rc=$?"
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err; echo "bench rc=$?"
tail -n 6 gpurun_out/r2d_split16.log gpurun_out/r2d_parity.log
grep -E "heat-maps|gradients|passed|failed|^E  " gpurun_out/r2d_sizes.log | head -40
head -n 30 gpurun_out/r2d_step_f16x3.md
python -c "
import json; d=json.load(open('gpurun_out/r2d_bench.json')); print(d['ms_per_step'], d['value'], d['e2e'], d['roofline']['achieved'], d['roofline']['conv_family'])"
