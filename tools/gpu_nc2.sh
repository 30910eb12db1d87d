cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_neg_cent_gpu.py -m gpu -q -x --timeout 300 2>&1 | tail -5
timeout 120 python tools/nc_bench.py 2>&1 | tail -5
python tools/nc_timeline.py gauss 2>&1 | grep "nc dbg" > gpurun_out/tl_gauss.txt; python tools/nc_timeline.py ota 2>&1 | grep "nc dbg" > gpurun_out/tl_ota.txt
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:nc_ --csv --log-file gpurun_out/nc_launches.csv python tools/nc_one.py gauss 3 > /dev/null 2>&1
grep -h "nc_" gpurun_out/nc_launches.csv | cut -d, -f5,19- | head -8
