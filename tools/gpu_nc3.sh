cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 300 $NCU -k regex:nc_v2_kernel -s 2 -c 1 -o gpurun_out/r02_nc_gauss python tools/nc_one.py gauss 4 > gpurun_out/ncu_nc_gauss.log 2>&1
timeout 300 $NCU -k regex:nc_prep_kernel -s 2 -c 1 -o gpurun_out/r02_nc_prep python tools/nc_one.py gauss 4 > gpurun_out/ncu_nc_prep.log 2>&1
