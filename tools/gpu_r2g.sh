#!/bin/bash
# what the driver runs at round end: full -m gpu suite, smoke, bench (both arms)
set +e
mkdir -p gpurun_out
timeout 2400 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider -s > gpurun_out/r2g_gpu_tests.log 2>&1; echo "gpu tests rc=$?"
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2g_smoke.log 2>&1; echo "smoke rc=$?"
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.err; echo "bench rc=$?"
timeout 900 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/r2g_bench_ref.json 2> gpurun_out/r2g_bench_ref.err; echo "bench ref rc=$?"
tail -n 5 gpurun_out/r2g_gpu_tests.log; grep -E "refiner train|C2|C5|gradients \(" gpurun_out/r2g_gpu_tests.log | head; cat gpurun_out/r2g_smoke.log | tail -2
python -c "
import json; d=json.load(open('gpurun_out/r2g_bench.json')); print(d['ms_per_step'], d['value'], d['e2e'], d['cpu_baseline'], d['clocks'])"
cat gpurun_out/r2g_bench_ref.json | head -c 800
