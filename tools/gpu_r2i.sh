#!/bin/bash
# 2 GPUs: data-parallel gradient semantics + the bench at N = 2 (stage-wise overlapped all-reduce in the graph)
set +e
mkdir -p gpurun_out
nvidia-smi -L
timeout 1500 python -m pytest tests/test_gpu_multi.py -q -p no:cacheprovider -s > gpurun_out/r2i_multi.log 2>&1; echo "multi rc=$?"
tail -n 6 gpurun_out/r2i_multi.log
timeout 600 python -m pytest tests/test_gpu_parity.py -q -p no:cacheprovider -s -k refiner_train_loop > gpurun_out/r2i_refiner.log 2>&1; echo "refiner rc=$?"
grep -E "refiner train|passed|failed|^E  " gpurun_out/r2i_refiner.log | head
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r2i_bench2.json 2> gpurun_out/r2i_bench2.err; echo "bench2 rc=$?"
tail -c 600 gpurun_out/r2i_bench2.err
python -c "
import json; d=json.loads(open('gpurun_out/r2i_bench2.json').read().strip().splitlines()[-1]); print(d['n_gpus'], d['ms_per_step'], d['value'], d['e2e'])"
timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2i_bench1.json 2> gpurun_out/r2i_bench1.err
python -c "
import json; d=json.loads(open('gpurun_out/r2i_bench1.json').read().strip().splitlines()[-1]); print(d['n_gpus'], d['ms_per_step'], d['value'], d['e2e'])"
