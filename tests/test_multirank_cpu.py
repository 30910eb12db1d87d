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


SHARD_WORKER = r"""
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, %(root)r)
import importlib.util
spec = importlib.util.spec_from_file_location("sharding", os.path.join(%(root)r, "aligner_b200", "sharding.py"))   # no CUDA library needed
sharding = importlib.util.module_from_spec(spec); spec.loader.exec_module(sharding)
from oracle import mas
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
rng = np.random.default_rng(7)                      # same seed on every rank: the plan is computed locally, identically
b, tx, ty = 24, 40, 120
t_x = rng.integers(5, tx + 1, b).astype(np.int32)
t_y = np.array([rng.integers(max(20, t_x[i]), ty + 1) for i in range(b)], np.int32)
values = rng.standard_normal((b, tx, ty)).astype(np.float32)
shards = sharding.balance_shards(t_x, t_y, world)
mine = shards[rank]
p = np.zeros((len(mine), tx, ty), np.int32)
mas.maximum_path_c_port(p, values[mine].copy(), t_x[mine].copy(), t_y[mine].copy())
dur = torch.zeros(b, tx, dtype=torch.int32)
dur[torch.from_numpy(mine)] = torch.from_numpy(p.sum(-1).astype(np.int32))
dist.all_reduce(dur, op=dist.ReduceOp.SUM)        # shards are disjoint: the sum assembles the whole batch
full = np.zeros((b, tx, ty), np.int32)
mas.maximum_path_c_port(full, values.copy(), t_x.copy(), t_y.copy())
assert np.array_equal(dur.numpy(), full.sum(-1)), "sharded durations differ from the unsharded run"
loads = sharding.shard_loads(t_x, t_y, shards)
assert sorted(np.concatenate(shards).tolist()) == list(range(b))
assert loads.max() - loads.min() <= sharding.item_cost(t_x, t_y).max()      # LPT bound
if rank == 0: print("SHARD_OK", loads.tolist())
dist.destroy_process_group()
"""


def test_two_rank_cost_balanced_shards(tmp_path):
    script = tmp_path / "shard_worker.py"
    script.write_text(SHARD_WORKER % {"root": str(ROOT)})
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                          "--master-port", "29534", str(script)], capture_output=True, text=True, timeout=240, env=env, cwd=str(ROOT))
    assert "SHARD_OK" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]


def test_balance_shards_properties():
    import importlib.util
    import numpy as np
    spec = importlib.util.spec_from_file_location("sharding", ROOT / "aligner_b200" / "sharding.py")
    sharding = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(sharding)
    rng = np.random.default_rng(3)
    for world in (1, 2, 3, 8):
        for b in (0, 1, 5, 257):
            t_x = rng.integers(1, 400, b)
            t_y = t_x + rng.integers(0, 1600, b)
            shards = sharding.balance_shards(t_x, t_y, world)
            assert len(shards) == world
            assert sorted(np.concatenate(shards).tolist() if b else []) == list(range(b))
            if b:
                loads = sharding.shard_loads(t_x, t_y, shards)
                assert loads.max() - loads.min() <= sharding.item_cost(t_x, t_y).max()
                for s in shards:        # descending cost inside a shard
                    c = sharding.item_cost(t_x, t_y)[s]
                    assert np.all(c[:-1] >= c[1:])
    order = sharding.lpt_order([3, 9, 9, 1], [10, 10, 10, 10])
    assert order.tolist() == [1, 2, 0, 3]


def test_reference_arm_runs_on_cpu():
    out = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1", "--workload", "c1"],
                         capture_output=True, text=True, timeout=240, cwd=str(ROOT))
    import json
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["value"] > 0 and line["cpu_baseline"]["kind"] in ("reference", "port")
    assert line["e2e"]["h2d_bytes_per_step"] == 0
