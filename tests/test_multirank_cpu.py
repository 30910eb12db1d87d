"""CPU, world_size 2 over gloo: the sharding / gather / neighbour cross-check logic bench.py uses at N > 1,
with the C port standing in for the GPU (no CUDA in this test)."""
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent

WORKER = r"""
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, %(root)r)
import bench
from oracle import mas
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
b, tx, ty = 6, 20, 50
def durations(seed):
    v, a, c = bench.make_batch(seed, b, tx, ty)
    p = np.zeros(v.shape, np.int32)
    mas.maximum_path_c_port(p, v.copy(), a, c)
    return torch.from_numpy(p.sum(-1).astype(np.int32))
mine = durations(1234 + 1 + rank)
gathered = [torch.empty_like(mine) for _ in range(world)]
dist.all_gather(gathered, mine)
nb = (rank + 1) %% world
ok = torch.tensor([int(torch.equal(durations(1234 + 1 + nb), gathered[nb]))])
dist.all_reduce(ok, op=dist.ReduceOp.MIN)
t = torch.tensor([float(rank + 1)], dtype=torch.float64)
dist.all_reduce(t, op=dist.ReduceOp.MAX)
assert ok.item() == 1 and t.item() == world
assert not torch.equal(gathered[0], gathered[1])          # different ranks really align different shards
if rank == 0: print("MULTIRANK_OK")
dist.destroy_process_group()
"""


def test_two_rank_gather_and_crosscheck(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER % {"root": str(ROOT)})
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                          "--master-port", "29533", str(script)], capture_output=True, text=True, timeout=240, env=env, cwd=str(ROOT))
    assert "MULTIRANK_OK" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]


def test_reference_arm_runs_on_cpu():
    out = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1", "--workload", "c1"],
                         capture_output=True, text=True, timeout=240, cwd=str(ROOT))
    import json
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["value"] > 0 and line["cpu_baseline"]["kind"] in ("reference", "port")
    assert line["e2e"]["h2d_bytes_per_step"] == 0
