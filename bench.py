#!/usr/bin/env python
"""bench.py -- MAS throughput on B200 (BASELINE.json metric), one JSON line on stdout.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c2]

Workload (config.workload): BASELINE.json configs[1], the VITS / Glow-TTS training-step
shape -- B=64 utterances, T_text=200, T_mel=1000, fp32 scores, every item full length.
A "step" is one monotonic_align.maximum_path(neg_cent, mask) call over that batch.

  value   cells/s with inputs resident in HBM (public device API, CUDA events, max over ranks)
  e2e     cells/s through the reference-facing host entry maximum_path_c(paths, values, t_xs, t_ys)
          with pinned HOST buffers: H2D of the scores and D2H of the result are inside the timed region
  roofline  the MAS kernel against measured HBM bandwidth, 8 algorithmic bytes per cell
  cpu_baseline  the reference's own core.pyx (oracle/_ref, -fopenmp) on this box's host cores

`--impl reference` times that CPU implementation alone, same metric/config.
Under torchrun every rank aligns its own batch (weak scaling, no collective on the hot path);
durations are all-gathered over NCCL after the timed region and cross-checked.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

WORKLOADS = {
    # name: (B, T_text, T_mel, description)
    "c1": (16, 100, 800, "BASELINE configs[0] LJSpeech-like"),
    "c2": (64, 200, 1000, "BASELINE configs[1] VITS/Glow-TTS training-step shape"),
    "c3": (32, 300, 1500, "BASELINE configs[2] OTA-style shape"),
    "c4": (8, 1000, 6000, "BASELINE configs[3] long-form"),
}
BYTES_PER_CELL = 8.0          # 4 B fp32 score read + 4 B fp32 path write
L2_BYTES = 126 << 20


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index: int):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nv = None

    def _once(self):
        nv = self.nv
        self.samples.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
        try:
            r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
        except Exception:
            r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
        names = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
                 0x80: "hw_power_brake", 0x2: "applications_clocks_setting"}
        for bit, name in names.items():
            if r & bit:
                self.reasons.add(name)

    def _loop(self):
        while not self._stop.is_set():
            try:
                self._once()
            except Exception:
                pass
            time.sleep(0.002)

    def __enter__(self):
        if self.nv is not None:
            self._thread = threading.Thread(target=self._loop, daemon=True)
            self._thread.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._thread is not None:
            self._thread.join()
            try:
                self._once()
            except Exception:
                pass

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unavailable"]}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


def make_batch(seed: int, b: int, tx: int, ty: int):
    rng = np.random.default_rng(seed)
    values = rng.standard_normal((b, tx, ty), dtype=np.float32)
    return values, np.full(b, tx, np.int32), np.full(b, ty, np.int32)


# --------------------------------------------------------------------------- CPU reference
def use_all_host_threads() -> int:
    """torchrun exports OMP_NUM_THREADS=1; the reference arm is entitled to every host core."""
    n = os.cpu_count() or 1
    os.environ["OMP_NUM_THREADS"] = str(n)
    try:
        import ctypes
        ctypes.CDLL("libgomp.so.1").omp_set_num_threads(n)
    except OSError:
        pass
    return n


def load_cpu_reference():
    """oracle/_ref (the reference's own core.pyx, -fopenmp) when present, else the C port."""
    use_all_host_threads()
    from oracle import mas
    ref = mas.load_reference_core("omp")
    if ref is not None:
        return "reference", ref.maximum_path_c, os.cpu_count()
    return "port", (lambda p, v, a, c: mas.maximum_path_c_port(p, v, a, c, omp=True)), mas.port_threads(True)


def time_cpu(fn, values, t_x, t_y, steps: int, warmup: int):
    paths = np.zeros(values.shape, np.int32)
    times = []
    for i in range(warmup + steps):
        v = values.copy()            # the reference clobbers its input (core.pyx:30)
        paths.fill(0)
        t0 = time.perf_counter()
        fn(paths, v, t_x, t_y)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    return times, paths


def run_reference(args, rank: int):
    if rank != 0:
        return
    b, tx, ty, desc = WORKLOADS[args.workload]
    kind, fn, cores = load_cpu_reference()
    values, t_x, t_y = make_batch(1234 + 1, b, tx, ty)
    times, _ = time_cpu(fn, values, t_x, t_y, args.steps, max(args.warmup, 1))
    cells = float(b) * tx * ty
    mean = float(np.mean(times))
    val = cells / mean
    line = {
        "impl": "reference", "metric": "mas_cells_per_sec", "value": val, "unit": "cells/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": mean * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "%s: B=%d T_text=%d T_mel=%d fp32, full lengths (%s)" % (args.workload, b, tx, ty, desc)},
        "utterances_per_sec": b / mean,
        "cpu_baseline": {"value": val, "unit": "cells/s", "cores": cores, "kind": kind,
                         "sample": "whole batch per step, maximum_path_c only (pre-zeroed paths), %d steps" % args.steps},
        "e2e": {"value": val, "unit": "cells/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------- ours
def run_ours(args, rank: int, world: int, local_rank: int):
    import torch
    import torch.distributed as dist

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    import aligner_b200.monotonic_align as ma
    from aligner_b200 import _lib
    from aligner_b200.monotonic_align.monotonic_align.core import maximum_path_c

    b, tx, ty, desc = WORKLOADS[args.workload]
    cells = float(b) * tx * ty
    per_set = int(cells) * 8
    nsets = max(3, int(np.ceil(3.0 * L2_BYTES / per_set)) + 1)       # rotate so a set is long gone from L2 when reused
    values_np, t_x, t_y = make_batch(1234 + 1 + rank, b, tx, ty)
    g = torch.Generator(device=dev).manual_seed(1234 + 1 + rank)
    vals = [torch.from_numpy(values_np).to(dev)] + [torch.randn(b, tx, ty, generator=g, device=dev) for _ in range(nsets - 1)]
    mask = torch.ones(b, tx, ty, device=dev)
    keep = [None] * nsets

    def step(i):
        keep[i % nsets] = ma.maximum_path(vals[i % nsets], mask)    # keeping nsets outputs alive rotates the output blocks too

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(max(args.warmup, 3)):
        step(i)
    barrier()
    # The K timed steps are K kernel launches of ~50 us each; issued from Python they are launch-bound on the host, so they are
    # captured once into a CUDA graph (same calls, same rotating buffers) and the replay is what is timed.
    graph, launches = None, 0
    if not args.no_graph:
        try:
            side = torch.cuda.Stream(dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):
                step(0)                                              # this stream's workspace must exist before capture
            torch.cuda.current_stream(dev).wait_stream(side)
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            n0 = _lib.launch_count()
            with torch.cuda.graph(graph, stream=side):
                for i in range(args.steps):
                    step(i)
            launches = _lib.launch_count() - n0
            graph.replay()                                           # warm replay
            torch.cuda.synchronize()
        except Exception as exc:                                     # capture not possible: time the eager loop
            print("cuda graph capture failed, timing eager launches:", exc, file=sys.stderr)
            graph = None
    barrier()
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clk:
        if graph is not None:
            start.record()
            graph.replay()
            end.record()
        else:
            n0 = _lib.launch_count()
            start.record()
            for i in range(args.steps):
                step(i)
            end.record()
            launches = _lib.launch_count() - n0
        torch.cuda.synchronize()
    ms = start.elapsed_time(end)
    barrier()
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    ms_step = ms_max / args.steps
    value = cells * world / (ms_step * 1e-3)

    # ---- kernel-only duration with events around single launches (graph-free, CPU overhead excluded)
    kt = []
    for i in range(min(args.steps, 50)):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); step(i); e1.record()
        kt.append((e0, e1))
    torch.cuda.synchronize()
    kernel_ms = float(np.median([a.elapsed_time(b_) for a, b_ in kt]))

    # ---- end to end through the host entry (pinned host buffers, copies inside the timed region)
    hv = torch.from_numpy(values_np).pin_memory()
    hp = torch.zeros(b, tx, ty, dtype=torch.int32).pin_memory()
    hv_np, hp_np = hv.numpy(), hp.numpy()
    e2e_times = []
    e2e_steps = max(3, min(args.steps, 20))
    for i in range(2 + e2e_steps):
        hp_np.fill(0)                                                # caller pre-zeroes, untimed (reference contract __init__.py:15)
        barrier()
        t0 = time.perf_counter()
        maximum_path_c(hp_np, hv_np, t_x, t_y)
        dt = time.perf_counter() - t0
        if i >= 2:
            e2e_times.append(dt)
    h2d, d2h = _lib.last_transfer_bytes()
    te = torch.tensor([float(np.mean(e2e_times))], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_val = cells * world / float(te.item())

    # ---- verification outside the timed region: durations all-gathered over NCCL, neighbour's shard re-run here
    out = ma.maximum_path_lengths(vals[0], torch.from_numpy(t_x).to(dev), torch.from_numpy(t_y).to(dev), dense=False, return_durations=True)
    dur = out["durations"]
    verified = None
    if world > 1:
        gathered = [torch.empty_like(dur) for _ in range(world)]
        dist.all_gather(gathered, dur)
        nb = (rank + 1) % world
        nv, nx, ny = make_batch(1234 + 1 + nb, b, tx, ty)
        mine = ma.maximum_path_lengths(torch.from_numpy(nv).to(dev), torch.from_numpy(nx).to(dev), torch.from_numpy(ny).to(dev),
                                       dense=False, return_durations=True)["durations"]
        ok = torch.tensor([int(torch.equal(mine, gathered[nb]))], device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        verified = bool(ok.item())
    host_ok = bool(np.array_equal(hp_np.sum(-1), dur.cpu().numpy()))

    if rank == 0:
        peak, peak_src = peaks()
        achieved = BYTES_PER_CELL * cells / (kernel_ms * 1e-3) / 1e9
        traffic = None
        tj = ROOT / "profiles" / "traffic.json"
        if tj.exists():
            traffic = json.loads(tj.read_text()).get(args.workload)
        cpu = None
        if world == 1 and not args.no_cpu:
            kind, fn, cores = load_cpu_reference()
            reps = 8
            times, ref_paths = time_cpu(fn, values_np, t_x, t_y, reps, 2)
            parity = bool(np.array_equal(ref_paths, hp_np))
            cpu = {"value": cells / float(np.mean(times)), "unit": "cells/s", "cores": cores, "kind": kind,
                   "sample": "the whole %s batch, %d repetitions of maximum_path_c with -fopenmp (%.1f ms each)" % (args.workload, reps, 1e3 * float(np.mean(times))),
                   "paths_equal_gpu": parity}
        line = {
            "metric": "mas_cells_per_sec", "value": value, "unit": "cells/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "%s: B=%d T_text=%d T_mel=%d fp32 per GPU, full lengths (%s)" % (args.workload, b, tx, ty, desc),
                       "api": "monotonic_align.maximum_path(value, mask) on CUDA tensors, dense fp32 path out",
                       "l2": "rotating %d input/output sets (%.0f MB) > 3x L2, no flush" % (nsets, nsets * per_set / 1e6),
                       "launch": "K steps replayed from one CUDA graph" if graph is not None else "eager launches"},
            "utterances_per_sec": b * world / (ms_step * 1e-3),
            "clocks": clk.summary(),
            "e2e": {"value": e2e_val, "unit": "cells/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "api": "maximum_path_c(paths, values, t_xs, t_ys) with pinned host numpy buffers", "ms_per_step": float(te.item()) * 1e3,
                    "paths_match_device_api": host_ok},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                         "kernel_ms": kernel_ms, "peak_source": peak_src, "algorithmic_bytes_per_launch": BYTES_PER_CELL * cells},
            "cpu_baseline": cpu,
            "durations_allgather_verified": verified,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-graph", action="store_true", help="time eager launches instead of a CUDA-graph replay")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if world != args.gpus and world == 1 and args.gpus > 1:
        # launched without torchrun: re-exec under it
        import subprocess
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus),
               "--master-addr", "127.0.0.1", "--master-port", "29511", __file__] + sys.argv[1:]
        raise SystemExit(subprocess.call(cmd))
    run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
