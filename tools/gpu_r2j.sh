#!/bin/bash
# compute-sanitizer over the tcgen05 / TMA kernel tests + determinism + refiner + bench
set +e
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
SEL='fprop_vs_emulation and (3x3_64_64 or 1x1_s2 or deconv4_ragged or final_24) or dgrad_vs_emulation and (3x3_s2_ragged and 1) or wgrad_vs_emulation and (3x3_128_ragged or 1x1_64_256 or deconv4_ragged)'
for tool in memcheck racecheck synccheck; do
  timeout 1500 $CS --tool $tool --print-limit 20 python -m pytest tests/test_gpu_split16.py -q -p no:cacheprovider -x -k "$SEL" > gpurun_out/r2j_sanitizer_$tool.log 2>&1
  echo "$tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" gpurun_out/r2j_sanitizer_$tool.log | tail -3
done
# the 3xTF32 pair kernel (remote mbarrier arrive, tune bit 3) under racecheck / memcheck
timeout 1500 $CS --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_parity.py -q -p no:cacheprovider -x -k "conv_family_vs_torch and 3-cfg1" > gpurun_out/r2j_sanitizer_tf32_memcheck.log 2>&1
grep -E "ERROR SUMMARY|passed|failed" gpurun_out/r2j_sanitizer_tf32_memcheck.log | tail -2
T="timeout 1200 python -m pytest -q -p no:cacheprovider -s"
$T tests/test_gpu_sizes.py -k "deterministic" > gpurun_out/r2j_det.log 2>&1; echo "determinism rc=$?"; tail -n 3 gpurun_out/r2j_det.log
$T tests/test_gpu_parity.py -k "refiner_train_loop or network_vs_reference_golden" > gpurun_out/r2j_ref.log 2>&1; echo "refiner+golden rc=$?"; grep -E "refiner train|passed|failed|^E  " gpurun_out/r2j_ref.log | head -5
timeout 600 python tools/step_table.py 32 f16x3 > gpurun_out/r2j_step_f16x3.md 2>/dev/null; head -n 30 gpurun_out/r2j_step_f16x3.md
timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2j_bench.json 2> gpurun_out/r2j_bench.err
python -c "
import json; d=json.loads(open('gpurun_out/r2j_bench.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'], d['e2e'], d['roofline']['conv_family']['per_kernel_ms_per_step'])"
