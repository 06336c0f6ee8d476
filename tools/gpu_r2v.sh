#!/bin/bash
# round 2 final evidence pass (1 GPU): sanitizer over the new kernels, launch list of one step, C5 line, both bench arms
set +e
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
SEL='bn_bwd_fused_entry or softargmax_bwd_split or finalize_scale or elementwise'
for tool in memcheck synccheck; do
  timeout 1200 $CS --tool $tool --print-limit 20 python -m pytest tests/test_gpu_split16.py -q -p no:cacheprovider -x -k "$SEL" > gpurun_out/r2v_sanitizer_$tool.log 2>&1
  echo "$tool rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/r2v_sanitizer_$tool.log | tail -2
done
timeout 1200 $CS --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_parity.py -q -p no:cacheprovider -x -k "softargmax_any_volume or input_pipeline or occluder" > gpurun_out/r2v_sanitizer_memcheck2.log 2>&1
grep -E "ERROR SUMMARY|passed|failed" gpurun_out/r2v_sanitizer_memcheck2.log | tail -2
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2v_launches.csv python tools/profile_step.py > gpurun_out/r2v_profile_step.log 2>&1
python tools/summarize_launches.py gpurun_out/r2v_launches.csv > gpurun_out/r2v_launches.md 2>&1; head -30 gpurun_out/r2v_launches.md
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --workload c5 > gpurun_out/r2v_bench_c5.json 2> gpurun_out/r2v_bench_c5.err
python -c "
import json; d=json.loads(open('gpurun_out/r2v_bench_c5.json').read().strip().splitlines()[-1]); print('c5', d['ms_per_step'], d['value'], d['roofline']['whole_step'])" || tail -3 gpurun_out/r2v_bench_c5.err
timeout 900 python bench.py --impl reference > gpurun_out/r2v_bench_ref.json 2> gpurun_out/r2v_bench_ref.err; echo "ref rc=$?"; cut -c1-400 gpurun_out/r2v_bench_ref.json
timeout 900 python bench.py > gpurun_out/r2v_bench.json 2> gpurun_out/r2v_bench.err; echo "bench rc=$?"; cat gpurun_out/r2v_bench.json
