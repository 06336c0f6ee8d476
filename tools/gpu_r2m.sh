#!/bin/bash
# round 2, call m: BN-backward two-pass reduction, fused finalize+scale, logit-gradient sink
mkdir -p gpurun_out
python -m pytest tests -q -x -m gpu > gpurun_out/r2m_gpu_tests.log 2>&1; echo "gpu tests rc=$?"
python tools/step_table.py > gpurun_out/r2m_step_f16x3.md 2> gpurun_out/r2m_step.err; echo "step rc=$?"
python bench.py > gpurun_out/r2m_bench.json 2> gpurun_out/r2m_bench.err; echo "bench rc=$?"
tail -5 gpurun_out/r2m_gpu_tests.log; sed -n 1,30p gpurun_out/r2m_step_f16x3.md; grep -A12 "elementwise pass" gpurun_out/r2m_step_f16x3.md
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2m_bench.json'))
print(d['ms_per_step'], d['value'], d['e2e'], d['roofline']['achieved'], d['roofline']['conv_family']['achieved'], d['gpu_launches'])
PY
tail -3 gpurun_out/r2m_bench.err
