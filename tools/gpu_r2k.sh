#!/bin/bash
# evidence pass: context lines (torch eager, other precisions, C5), ncu launch list of one step, full ncu of the dominant kernel
set +e
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -p no:cacheprovider -s -k refiner_train_loop > gpurun_out/r2k_refiner.log 2>&1; echo "refiner rc=$?"; grep -E "refiner train|passed|failed|^E  " gpurun_out/r2k_refiner.log | head -4
timeout 900 python tools/torch_eager_b200.py 128 10 > gpurun_out/r2k_torch_eager.md 2>&1; cat gpurun_out/r2k_torch_eager.md | tail -6
for prec in tf32 tf32x3; do
  timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --precision $prec > gpurun_out/r2k_bench_$prec.json 2>/dev/null
  python -c "
import json; d=json.loads(open('gpurun_out/r2k_bench_$prec.json').read().strip().splitlines()[-1]); print('$prec', d['ms_per_step'], d['value'])"
done
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --workload c5 > gpurun_out/r2k_bench_c5.json 2> gpurun_out/r2k_bench_c5.err
python -c "
import json; d=json.loads(open('gpurun_out/r2k_bench_c5.json').read().strip().splitlines()[-1]); print('c5', d['ms_per_step'], d['value'], d['config']['workload'][:60], d['roofline']['whole_step'])" || tail -3 gpurun_out/r2k_bench_c5.err
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2k_launches.csv python tools/profile_step.py > gpurun_out/r2k_profile_step.log 2>&1
python tools/summarize_launches.py gpurun_out/r2k_launches.csv > gpurun_out/r2k_launches.md 2>&1; head -40 gpurun_out/r2k_launches.md
timeout 300 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"wgrad16_kernel" -s 2 -c 1 -o gpurun_out/r2k_w16 -f python tools/one_conv16.py 128 16 16 256 256 3 1 wgrad > gpurun_out/r2k_ncu_w16.log 2>&1
python tools/ncu_summary.py gpurun_out/r2k_w16.ncu-rep 12 > gpurun_out/r2k_sum_w16.txt 2>&1; head -24 gpurun_out/r2k_sum_w16.txt
