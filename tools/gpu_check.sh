# tools/gpu_check.sh -- one gpurun call: GPU tests, then the bench (both arms).  Output under gpurun_out/.
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
tail -c 300 gpurun_out/bench_n1.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_n1.json").read().strip().splitlines()[-1])
print("value %.3e ms/step %.4f frac %.3f e2e %.3e" % (d["value"], d["ms_per_step"], d["roofline"]["frac"], d["e2e"]["value"]))
for c in d.get("configs") or []:
    print("  %-28s %9.3f ms  frac %.3f  ok=%s  %s" % (c["name"], c["ms"], c["roofline"]["frac"], c["paths_equal_oracle"], c["kernel"][:60]))
print(json.dumps(d.get("neg_cent"), indent=1)[:1500])
print(json.dumps(d.get("strong_scaling_c5"), indent=1)[:1500])
print(json.dumps(d.get("cpu_baseline"), indent=1)[:1800])
print(json.dumps(d.get("e2e_device_api"), indent=1))
PY
