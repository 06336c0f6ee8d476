#!/bin/bash
# 2 GPUs, final code: data-parallel gradient semantics, bench at N = 2 (weak and strong) and N = 1 on the same box, reference arm under torchrun
set +e
mkdir -p gpurun_out
nvidia-smi -L
timeout 1500 python -m pytest tests/test_gpu_multi.py -q -p no:cacheprovider -s > gpurun_out/r2w_multi.log 2>&1; echo "multi rc=$?"
tail -n 6 gpurun_out/r2w_multi.log
run() { timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu-baseline $2 > gpurun_out/$3.json 2> gpurun_out/$3.err; echo "$3 rc=$?"; python -c "
import json; d=json.loads(open('gpurun_out/$3.json').read().strip().splitlines()[-1]); print(d['n_gpus'], d['scaling'], d['ms_per_step'], d['value'], d['e2e'])" || tail -c 600 gpurun_out/$3.err; }
run 29711 "" r2w_bench2
run 29712 "--strong" r2w_bench2_strong
timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2w_bench1.json 2> gpurun_out/r2w_bench1.err
python -c "
import json; d=json.loads(open('gpurun_out/r2w_bench1.json').read().strip().splitlines()[-1]); print(d['n_gpus'], d['ms_per_step'], d['value'], d['e2e'])"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29713 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 > gpurun_out/r2w_ref2.json 2> gpurun_out/r2w_ref2.err; echo "ref2 rc=$?"; cut -c1-200 gpurun_out/r2w_ref2.json
