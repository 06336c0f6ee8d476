# ncu --set full captures of single bench layers (tools/one_conv.py), summarised with tools/ncu_summary.py.
#   usage (on the GPU box): bash tools/ncu_layers.sh <tag> "<one_conv args>" ["<one_conv args>" ...]
cd $GRAFT_REPO_ROOT
tag=$1; shift
i=0
for cfg in "$@"; do
  i=$((i+1))
  timeout 300 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
      -k regex:"conv_(fprop|wgrad)_tc" -s 2 -c 1 -o gpurun_out/${tag}_$i -f python tools/one_conv.py $cfg > gpurun_out/${tag}_ncu$i.log 2>&1
  echo "# one_conv.py $cfg" > gpurun_out/${tag}_sum$i.txt
  python tools/ncu_summary.py gpurun_out/${tag}_$i.ncu-rep 45 >> gpurun_out/${tag}_sum$i.txt 2>&1
done
