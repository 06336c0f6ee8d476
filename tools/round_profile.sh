# End-of-round evidence run (GPU box): parity tests, layer A/B, bench line, ncu launch list of one
# training step, ncu --set full captures of the dominant kernels.   usage: bash tools/round_profile.sh <tag>
cd $GRAFT_REPO_ROOT
tag=${1:-rX}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_gputests.log 2>&1; echo "tests rc $?" >> gpurun_out/${tag}_gputests.log
if [ -f build/variants/libepb_base.so ]; then
  EPB_LIB_PATH=build/variants/libepb_base.so timeout 200 python tools/layer_sweep.py > gpurun_out/${tag}_sweep_base.txt 2>&1
fi
timeout 200 python tools/layer_sweep.py > gpurun_out/${tag}_sweep_new.txt 2>&1
timeout 500 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc $?"
timeout 400 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/${tag}_launches.csv python tools/profile_step.py > gpurun_out/${tag}_profile_step.log 2>&1
python tools/summarize_launches.py gpurun_out/${tag}_launches.csv > gpurun_out/${tag}_launches.md 2>&1
bash tools/ncu_layers.sh ${tag}_full "128 16 16 256 256 3 1 fprop" "128 64 64 64 256 1 1 fprop" "128 64 64 256 1024 1 1 wgrad"
tail -3 gpurun_out/${tag}_gputests.log; cat gpurun_out/${tag}_bench.json
