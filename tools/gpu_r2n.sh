#!/bin/bash
# round 2, call n: BN backward per-kernel times (ncu launch list on single tensors), step table, bench
mkdir -p gpurun_out
python -m pytest tests/test_gpu_split16.py tests/test_gpu_sizes.py -q -x -k "bn_bwd or softargmax_bwd or finalize_scale or c3 or determin" > gpurun_out/r2n_tests.log 2>&1; echo "tests rc=$?"
: > gpurun_out/r2n_bn16.txt
for mc in "524288 256" "131072 512" "32768 1024" "32768 256" "8192 512"; do
  python tools/one_bn16.py $mc 7 >> gpurun_out/r2n_bn16.txt 2>&1
done
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2n_bn_launches.csv python tools/one_bn16.py 524288 256 1 > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2n_bn_launches_small.csv python tools/one_bn16.py 32768 256 1 > /dev/null 2>&1
python tools/step_table.py > gpurun_out/r2n_step_f16x3.md 2> gpurun_out/r2n_step.err; echo "step rc=$?"
python bench.py > gpurun_out/r2n_bench.json 2> gpurun_out/r2n_bench.err; echo "bench rc=$?"
tail -3 gpurun_out/r2n_tests.log; cat gpurun_out/r2n_bn16.txt
grep "bn_bwd\|softargmax\|colsum" gpurun_out/r2n_bn_launches.csv | awk -F'","' '{print $5, $NF}' | tail -12
grep "bn_bwd\|softargmax\|colsum" gpurun_out/r2n_bn_launches_small.csv | awk -F'","' '{print $5, $NF}' | tail -12
sed -n 1,22p gpurun_out/r2n_step_f16x3.md; grep -A16 "elementwise pass" gpurun_out/r2n_step_f16x3.md
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2n_bench.json'))
print(d['ms_per_step'], d['value'], d['e2e'], d['roofline']['achieved'], d['roofline']['conv_family']['achieved'], d['gpu_launches'])
PY
