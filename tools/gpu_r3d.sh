#!/bin/bash
set +e
mkdir -p gpurun_out
python tools/bench_aux.py > gpurun_out/r3d_aux.jsonl 2> gpurun_out/r3d_aux.err; echo "aux rc=$?"
cut -c1-190 gpurun_out/r3d_aux.jsonl; tail -3 gpurun_out/r3d_aux.err
