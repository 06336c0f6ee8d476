#!/bin/bash
set +e
mkdir -p gpurun_out
python -m pytest tests/test_gpu_split16.py tests/test_gpu_parity.py -q -x -k "elementwise or network_vs_reference_golden" > gpurun_out/r3e_unit.log 2>&1; echo "unit rc=$?"; tail -3 gpurun_out/r3e_unit.log
python tools/step_table.py 2>/dev/null | grep -E "^step|im2col|conv16_fprop\`|bn_bwd_split\`"
python bench.py --no-cpu-baseline > gpurun_out/r3e_bench.json 2> gpurun_out/r3e_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r3e_bench.json'))
print(d['ms_per_step'], d['value'], d['e2e']['value'], d['roofline']['non_conv_ms_per_step']['epb_im2col_split'])
PY
