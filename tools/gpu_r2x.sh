#!/bin/bash
# 8 GPUs: the bench line the driver's scaling run will ask for (weak scaling), shortest form
set +e
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29721 bench.py --gpus 8 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2x_bench8.json 2> gpurun_out/r2x_bench8.err; echo "bench8 rc=$?"
python -c "
import json; d=json.loads(open('gpurun_out/r2x_bench8.json').read().strip().splitlines()[-1]); print(d['n_gpus'], d['scaling'], d['ms_per_step'], d['value'], d['e2e'], d['clocks'])" || tail -c 1500 gpurun_out/r2x_bench8.err
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2x_bench1.json 2> gpurun_out/r2x_bench1.err
python -c "
import json; d=json.loads(open('gpurun_out/r2x_bench1.json').read().strip().splitlines()[-1]); print(d['n_gpus'], d['ms_per_step'], d['value'], d['e2e'])"
