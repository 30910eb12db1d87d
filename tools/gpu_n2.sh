cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node ${NG:-2} --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus ${NG:-2} --steps 20 --warmup 5 > gpurun_out/bench_n${NG:-2}.json 2> gpurun_out/bench_n${NG:-2}.err; echo "rc=$?"
tail -c 400 gpurun_out/bench_n${NG:-2}.err
NGV=${NG:-2} python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_n%s.json" % __import__("os").environ["NGV"]).read().strip().splitlines()[-1])
print("value %.3e ms/step %.4f e2e %.3e verified %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["durations_allgather_verified"]))
print(json.dumps(d["e2e"], indent=1)[:900])
print(json.dumps(d.get("strong_scaling_c5"), indent=1)[:1800])
PY
