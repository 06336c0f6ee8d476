#!/bin/bash
set +e
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -q -p no:cacheprovider > gpurun_out/r3h_multi.log 2>&1; echo "multi rc=$?"; tail -2 gpurun_out/r3h_multi.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29731 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r3h_bench2.json 2> gpurun_out/r3h_bench2.err; echo "bench2 rc=$?"
python -c "
import json; d=json.loads(open('gpurun_out/r3h_bench2.json').read().strip().splitlines()[-1]); print(d['n_gpus'], d['ms_per_step'], d['value'], d['e2e']['value'])" || tail -c 800 gpurun_out/r3h_bench2.err
