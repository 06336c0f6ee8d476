# second evidence run of the round: full GPU tests (in-tree lib), flat-producer variant parity + A/B,
# aux kernel throughput.   usage: bash tools/round_profile2.sh <tag>
cd $GRAFT_REPO_ROOT
tag=${1:-rX}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_gputests.log 2>&1; echo "tests rc $?" >> gpurun_out/${tag}_gputests.log
timeout 200 python tools/bench_aux.py > gpurun_out/${tag}_aux.jsonl 2> gpurun_out/${tag}_aux.err
timeout 200 python tools/layer_sweep.py > gpurun_out/${tag}_sweep_intree.txt 2>&1
EPB_LIB_PATH=build/variants/libepb_flat.so timeout 200 python tools/layer_sweep.py > gpurun_out/${tag}_sweep_flat.txt 2>&1
EPB_LIB_PATH=build/variants/libepb_flat.so timeout 300 python -m pytest tests/test_gpu_parity.py -q -x -k "conv_family or network or graphed or script_flow" > gpurun_out/${tag}_flat_tests.log 2>&1; echo "flat tests rc $?" >> gpurun_out/${tag}_flat_tests.log
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/${tag}_bench_intree.json 2> gpurun_out/${tag}_bench_intree.err
EPB_LIB_PATH=build/variants/libepb_flat.so timeout 300 python bench.py --no-cpu-baseline > gpurun_out/${tag}_bench_flat.json 2> gpurun_out/${tag}_bench_flat.err
tail -3 gpurun_out/${tag}_gputests.log; tail -2 gpurun_out/${tag}_flat_tests.log; cat gpurun_out/${tag}_aux.jsonl
paste gpurun_out/${tag}_sweep_intree.txt gpurun_out/${tag}_sweep_flat.txt | cut -c1-150
python -c "
import json
for n in ('intree','flat'):
    try:
        d=json.load(open('gpurun_out/${tag}_bench_%s.json'%n)); print(n, d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['conv_family']['per_kernel_ms_per_step'])
    except Exception as e: print(n,'ERR',e)
"
