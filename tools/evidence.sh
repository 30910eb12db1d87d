# tools/evidence.sh -- one gpurun call that regenerates the round's profiles (copied into profiles/ by hand afterwards)
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
R=r02
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 300 $NCU -k regex:mas_kernel -s 2 -c 1 -o gpurun_out/${R}_mas_c2 python tools/run_one.py c2 4 > gpurun_out/ncu_c2.log 2>&1
timeout 300 $NCU -k regex:mas_kernel -s 2 -c 1 -o gpurun_out/${R}_mas_c5b python tools/run_one.py c5b 4 > gpurun_out/ncu_c5b.log 2>&1
timeout 300 $NCU -k regex:nc_v2_kernel -s 2 -c 1 -o gpurun_out/${R}_nc_gauss python tools/nc_one.py gauss 4 > gpurun_out/ncu_nc_gauss.log 2>&1
timeout 300 $NCU -k regex:nc_v2_kernel -s 2 -c 1 -o gpurun_out/${R}_nc_ota python tools/nc_one.py ota 4 > gpurun_out/ncu_nc_ota.log 2>&1
timeout 300 $NCU -k regex:nc_prep_kernel -s 2 -c 1 -o gpurun_out/${R}_nc_prep python tools/nc_one.py gauss 4 > gpurun_out/ncu_nc_prep.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${R}_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu --headline-only > gpurun_out/bench_under_ncu.log 2>&1
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:nc_ --csv --log-file gpurun_out/${R}_launches_nc.csv python tools/nc_one.py gauss 3 > /dev/null 2>&1
for w in c1 c2 c3 c4 c5a c5b c5c; do timeout 200 python tools/mas_sweep.py $w 2>&1 | grep None | cut -c1-240; done > gpurun_out/${R}_shape_defaults.txt
timeout 120 python tools/nc_bench.py > gpurun_out/${R}_nc_bench.txt 2>&1
timeout 200 python tools/fused_bench.py > gpurun_out/${R}_fused_bench.txt 2>&1
for n in ${R}_mas_c2 ${R}_mas_c5b ${R}_nc_gauss ${R}_nc_ota ${R}_nc_prep; do python tools/ncu_summary.py gpurun_out/$n.ncu-rep > gpurun_out/$n.summary.txt 2>&1; done
python tools/prof_lines.py gpurun_out/${R}_mas_c2.ncu-rep > gpurun_out/${R}_mas_c2.lines.txt 2>&1 || true
python tools/prof_lines.py gpurun_out/${R}_nc_gauss.ncu-rep > gpurun_out/${R}_nc_gauss.lines.txt 2>&1 || true
rm -f gpurun_out/*.ncu-rep
cat gpurun_out/${R}_shape_defaults.txt gpurun_out/${R}_nc_bench.txt gpurun_out/${R}_fused_bench.txt
