set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 300 $NCU -k regex:mas_kernel -s 2 -c 1 -o gpurun_out/r01_mas_c2 python tools/run_one.py c2 4 > gpurun_out/ncu_c2.log 2>&1
timeout 300 $NCU -k regex:mas_kernel -s 2 -c 1 -o gpurun_out/r01_mas_c5b python tools/run_one.py c5b 4 > gpurun_out/ncu_c5b.log 2>&1
timeout 300 $NCU -k regex:mas_kernel -s 2 -c 1 -o gpurun_out/r01_mas_c4 python tools/run_one.py c4 4 > gpurun_out/ncu_c4.log 2>&1
timeout 300 $NCU -k regex:nc_tc_kernel -s 3 -c 1 -o gpurun_out/r01_nc_gauss_tc python tools/nc_bench.py > gpurun_out/ncu_nc.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/bench_under_ncu.log 2>&1
for w in c1 c2 c3 c4 c5a c5b c5c; do timeout 200 python tools/mas_sweep.py $w 2>&1 | grep None | cut -c1-240; done > gpurun_out/shape_defaults.txt
timeout 120 python tools/nc_bench.py > gpurun_out/nc_bench.txt 2>&1
timeout 300 python bench.py --steps 200 --warmup 10 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
for n in r01_mas_c2 r01_mas_c5b r01_mas_c4 r01_nc_gauss_tc; do python tools/ncu_summary.py gpurun_out/$n.ncu-rep > gpurun_out/$n.summary.txt 2>&1; done
python tools/prof_lines.py gpurun_out/r01_mas_c2.ncu-rep > gpurun_out/r01_mas_c2.lines.txt 2>&1 || true
rm -f gpurun_out/*.ncu-rep
ls -la gpurun_out | head -30
cat gpurun_out/shape_defaults.txt gpurun_out/nc_bench.txt
tail -c 600 gpurun_out/bench_n1.json; tail -c 400 gpurun_out/bench_ref.json
