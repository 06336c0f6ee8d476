# A-stream probes (EPB_TUNE bits: 0-1 prefetch flavour, 2 evict-first demand loads, 3 CTA-scope pair arrive)
cd $GRAFT_REPO_ROOT
for cfg in "128 64 64 256 1024 1 1 dgrad" "128 32 32 512 128 1 1 fprop" "128 64 64 256 64 1 1 fprop" "128 32 32 128 128 3 1 fprop" "128 16 16 1024 256 1 1 fprop"; do
  for t in 0 1 2 3 5 6 9 13 14; do
    echo -n "tune=$t pair=1 | "; EPB_TUNE=$t timeout 100 python tools/one_conv.py $cfg 2>&1 | tail -1
  done
  echo -n "tune=1 pair=0 | "; EPB_CTA_PAIR=0 EPB_TUNE=1 timeout 100 python tools/one_conv.py $cfg 2>&1 | tail -1
  echo -n "tune=6 pair=0 | "; EPB_CTA_PAIR=0 EPB_TUNE=6 timeout 100 python tools/one_conv.py $cfg 2>&1 | tail -1
done
