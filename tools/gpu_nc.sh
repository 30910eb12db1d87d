cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_neg_cent_gpu.py -m gpu -q -x --timeout 300 2>&1 | tail -5
timeout 120 python tools/nc_bench.py 2>&1 | tail -5
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 300 $NCU -k regex:nc_v2_kernel -s 2 -c 1 -o gpurun_out/r02_nc_gauss python tools/nc_one.py gauss 4 > gpurun_out/ncu_nc_gauss.log 2>&1
timeout 300 $NCU -k regex:nc_v2_kernel -s 2 -c 1 -o gpurun_out/r02_nc_ota python tools/nc_one.py ota 4 > gpurun_out/ncu_nc_ota.log 2>&1
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:nc_ --csv --log-file gpurun_out/nc_launches.csv python tools/nc_one.py gauss 3 > /dev/null 2>&1
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:nc_ --csv --log-file gpurun_out/nc_launches_ota.csv python tools/nc_one.py ota 3 > /dev/null 2>&1
grep -h "nc_" gpurun_out/nc_launches.csv gpurun_out/nc_launches_ota.csv | cut -d, -f5,15- | head -20
