cd $GRAFT_REPO_ROOT
for w in ${WL:-x5 x6 c5a c5c x8 x9 z1 z3 z2 z4 q4 q3 x4}; do
  for m in 0 1000000; do
    echo -n "latmax=$m "; ALB200_LATENCY_MAX_B=$m timeout 120 python tools/mas_sweep.py $w 2>&1 | grep "None" | cut -c1-150
  done
done
