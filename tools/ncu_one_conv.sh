set -x
cd $GRAFT_REPO_ROOT
for cfg in "128 64 64 256 1024 1 1 dgrad" "128 32 32 512 128 1 1 fprop" "128 64 64 256 1024 1 1 wgrad" "128 16 16 256 256 3 1 fprop"; do
  timeout 120 python tools/one_conv.py $cfg 2>&1 | tail -1
done
i=0
for cfg in "128 64 64 256 1024 1 1 dgrad" "128 32 32 512 128 1 1 fprop" "128 64 64 256 1024 1 1 wgrad"; do
  i=$((i+1))
  timeout 300 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"conv_(fprop|wgrad)_tc" -s 2 -c 1 -o gpurun_out/r29_one$i -f python tools/one_conv.py $cfg > gpurun_out/r29_ncu$i.log 2>&1
  python tools/ncu_summary.py gpurun_out/r29_one$i.ncu-rep 14 > gpurun_out/r29_sum$i.txt 2>&1
done
