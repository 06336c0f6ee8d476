#!/bin/bash
# round 2, call l: per-shape elementwise table, aux kernels, new soft-argmax shapes
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -q -x -k "softargmax" > gpurun_out/r2l_softargmax.log 2>&1; echo "softargmax rc=$?"
python tools/step_table.py > gpurun_out/r2l_step_f16x3.md 2> gpurun_out/r2l_step.err; echo "step rc=$?"
python tools/bench_aux.py > gpurun_out/r2l_aux.jsonl 2> gpurun_out/r2l_aux.err; echo "aux rc=$?"
for mc in "524288 64" "524288 256" "131072 128" "131072 512" "32768 256" "32768 1024" "8192 512" "8192 2048"; do
  python tools/one_bn16.py $mc 7 >> gpurun_out/r2l_bn16.txt 2>&1
done
tail -3 gpurun_out/r2l_softargmax.log; tail -30 gpurun_out/r2l_step_f16x3.md; cat gpurun_out/r2l_aux.jsonl | cut -c1-200; tail -5 gpurun_out/r2l_aux.err; cat gpurun_out/r2l_bn16.txt
