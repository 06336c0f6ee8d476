# last evidence run of a round (GPU box), most important first: parity tests, bench line, smoke(), the C5 sanity
# check, ncu launch list of one step.   usage: bash tools/round_final.sh <tag>
cd $GRAFT_REPO_ROOT
tag=${1:-rX}
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_gputests.log 2>&1; echo "tests rc $?" >> gpurun_out/${tag}_gputests.log
timeout 400 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc $?"
timeout 120 python __graft_entry__.py --smoke > gpurun_out/${tag}_smoke.log 2>&1; echo "smoke rc $?" >> gpurun_out/${tag}_smoke.log
timeout 200 python tools/c5_check.py 2 > gpurun_out/${tag}_c5.log 2>&1; echo "c5 rc $?" >> gpurun_out/${tag}_c5.log
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/${tag}_launches.csv python tools/profile_step.py > gpurun_out/${tag}_profile_step.log 2>&1
python tools/summarize_launches.py gpurun_out/${tag}_launches.csv > gpurun_out/${tag}_launches.md 2>&1
tail -3 gpurun_out/${tag}_gputests.log; tail -2 gpurun_out/${tag}_smoke.log; tail -3 gpurun_out/${tag}_c5.log; cat gpurun_out/${tag}_bench.json
