cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
bash tools/san_all.sh 2>&1 | grep -v "^ *[0-9]* =========" | tail -8
for w in c1 c2 c3 c4 c5a c5b c5c; do timeout 200 python tools/mas_sweep.py $w 2>&1 | grep None | cut -c1-240; done > gpurun_out/r02_shape_defaults.txt
cat gpurun_out/r02_shape_defaults.txt | cut -c1-110
