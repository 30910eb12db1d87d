#!/usr/bin/env python
"""bench.py -- MAS throughput on B200 (BASELINE.json metric), one JSON line on stdout.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c2] [--headline-only]

Headline workload (config.workload): BASELINE.json configs[1], the VITS / Glow-TTS training-step
shape -- B=64 utterances, T_text=200, T_mel=1000, fp32 scores, every item full length.
A "step" is one monotonic_align.maximum_path(neg_cent, mask) call over that batch.

  value         cells/s with inputs resident in HBM (public device API, CUDA events, max over ranks)
  e2e           cells/s through the reference-facing host entry maximum_path_c(paths, values, t_xs, t_ys)
                with pinned HOST buffers: H2D of the scores and D2H of the result are inside the timed region
  roofline      the MAS kernel against measured HBM bandwidth, 8 algorithmic bytes per cell
  cpu_baseline  the reference's own core.pyx (oracle/_ref) on this box's host cores: -fopenmp on all cores,
                serial as the reference ships it (setup.py:5-9), and its Python API fed CUDA tensors
  configs       every other BASELINE.json configuration, device-timed the same way (N = 1 only):
                c1, c3, c4, c5 (B=2048 mixed lengths) and 4096x200x1000, each with its roofline fraction and a
                parity check of a sample against the CPU oracle
  neg_cent      the score-matrix kernels (Gaussian C2, OTA C3) and the neg_cent -> MAS pipeline (N = 1 only)
  strong_scaling_c5   under --gpus N: ONE B=8192 mixed-length batch (BASELINE configs[4]) split across the ranks
                with aligner_b200.sharding.balance_shards, processed in chunks of <= 16 GB per GPU; per-rank
                time, load imbalance, durations all-gathered over NCCL and cross-checked

`--impl reference` times the reference's CPU implementation alone, same metric/config.
Under torchrun every rank aligns its own headline batch (weak scaling, no collective on the hot path).
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
C5_TX, C5_TY = 400, 2000          # BASELINE configs[4]: t_x ~ U{50..400}, t_y ~ U{max(200, t_x)..2000}
BYTES_PER_CELL = 8.0              # 4 B fp32 score read + 4 B fp32 path write
L2_BYTES = 126 << 20


def workload_string(name: str) -> str:
    """One string for both arms (the driver compares them)."""
    b, tx, ty, desc = WORKLOADS[name]
    return "%s: B=%d T_text=%d T_mel=%d fp32, full lengths (%s)" % (name, b, tx, ty, desc)


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return float(d["hbm_gbs"]), float(d.get("bf16_tflops", 0.0)), "measured (MEASURED_PEAKS.json)"
    return 6650.0, 1500.0, "fallback (B200_PROFILING.md)"


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


def c5_lengths(b: int, seed: int = 1239):
    """BASELINE configs[4] / SURVEY.md 8d: t_x ~ U{50..400}, t_y ~ U{max(200, t_x)..2000}."""
    rng = np.random.default_rng(seed)
    t_x = rng.integers(50, C5_TX + 1, b).astype(np.int32)
    t_y = np.array([rng.integers(max(200, t_x[i]), C5_TY + 1) for i in range(b)], np.int32)
    return t_x, t_y


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


def load_cpu_reference(kind: str = "omp"):
    """oracle/_ref (the reference's own core.pyx; 'omp' = -fopenmp, 'serial' = as shipped) when present, else the C port."""
    use_all_host_threads()
    from oracle import mas
    ref = mas.load_reference_core(kind)
    if ref is not None:
        return "reference", ref.maximum_path_c, (os.cpu_count() if kind == "omp" else 1)
    omp = kind == "omp"
    return "port", (lambda p, v, a, c: mas.maximum_path_c_port(p, v, a, c, omp=omp)), (mas.port_threads(True) if omp else 1)


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


def reference_api(core_fn):
    """The reference's Python API (monotonic_align/__init__.py:6-21) bound to its own compiled core.  The reference tree is
    not on the GPU box, so its 11 staging lines are restated here one for one (line numbers on the right)."""
    import torch

    def maximum_path(value, mask):
        value = value * mask                                            # __init__.py:11
        device, dtype = value.device, value.dtype                       # __init__.py:12-13
        value = value.data.cpu().numpy().astype(np.float32)             # __init__.py:14   (device -> host, blocking)
        path = np.zeros_like(value).astype(np.int32)                    # __init__.py:15
        mask = mask.data.cpu().numpy()                                  # __init__.py:16   (device -> host)
        t_x_max = mask.sum(1)[:, 0].astype(np.int32)                    # __init__.py:18
        t_y_max = mask.sum(2)[:, 0].astype(np.int32)                    # __init__.py:19
        core_fn(path, value, t_x_max, t_y_max)                          # __init__.py:20
        return torch.from_numpy(path).to(device=device, dtype=dtype)    # __init__.py:21   (host -> device)
    return maximum_path


def run_reference(args, rank: int):
    if rank != 0:
        return
    b, tx, ty, desc = WORKLOADS[args.workload]
    kind, fn, cores = load_cpu_reference("omp")
    values, t_x, t_y = make_batch(1234 + 1, b, tx, ty)
    times, _ = time_cpu(fn, values, t_x, t_y, args.steps, max(args.warmup, 1))
    cells = float(b) * tx * ty
    mean = float(np.mean(times))
    val = cells / mean
    line = {
        "impl": "reference", "metric": "mas_cells_per_sec", "value": val, "unit": "cells/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": mean * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_string(args.workload)},
        "utterances_per_sec": b / mean,
        "cpu_baseline": {"value": val, "unit": "cells/s", "cores": cores, "kind": kind,
                         "sample": "whole batch per step, maximum_path_c only (pre-zeroed paths), -fopenmp on every host core, %d steps" % args.steps,
                         "best_ms": float(np.min(times)) * 1e3, "median_ms": float(np.median(times)) * 1e3},
        "e2e": {"value": val, "unit": "cells/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------- device timing helpers
class DeviceTimer:
    """K steps captured into one CUDA graph and replayed (the launches are ~30-1000 us each; issued from Python they would be
    launch-bound on the host), timed with CUDA events on the replay stream.  Falls back to eager launches."""

    def __init__(self, torch, dev, lib, use_graph=True):
        self.torch, self.dev, self.lib, self.use_graph = torch, dev, lib, use_graph

    def run(self, step, k: int, warmup: int = 3, sampler=None):
        torch, dev = self.torch, self.dev
        for i in range(warmup):
            step(i)
        torch.cuda.synchronize()
        graph, launches = None, 0
        if self.use_graph:
            try:
                side = torch.cuda.Stream(dev)
                side.wait_stream(torch.cuda.current_stream(dev))
                with torch.cuda.stream(side):
                    step(0)                                              # this stream's workspace must exist before capture
                torch.cuda.current_stream(dev).wait_stream(side)
                torch.cuda.synchronize()
                graph = torch.cuda.CUDAGraph()
                n0 = self.lib.launch_count()
                with torch.cuda.graph(graph, stream=side):
                    for i in range(k):
                        step(i)
                launches = self.lib.launch_count() - n0
                graph.replay()                                           # warm replay
                torch.cuda.synchronize()
            except Exception as exc:                                     # capture not possible: time the eager loop
                print("cuda graph capture failed, timing eager launches:", exc, file=sys.stderr)
                graph = None
        start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ctx = sampler if sampler is not None else _Null()
        with ctx:
            if graph is not None:
                start.record()
                graph.replay()
                end.record()
            else:
                n0 = self.lib.launch_count()
                start.record()
                for i in range(k):
                    step(i)
                end.record()
                launches = self.lib.launch_count() - n0
            torch.cuda.synchronize()
        ms = start.elapsed_time(end)
        how = "K steps replayed from one CUDA graph" if graph is not None else "eager launches"
        del graph
        return ms / k, int(launches), how


class _Null:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


def rotating_sets(per_set_bytes: int, lo: int = 2, hi: int = 8) -> int:
    """How many input/output sets to rotate through so that a set is long gone from the 126 MB L2 when it is reused."""
    return int(min(max(lo, int(np.ceil(3.0 * L2_BYTES / per_set_bytes)) + 1), hi))


def oracle_sample_ok(values_t, t_x, t_y, result, idx, field="durations"):
    """Parity of a sample of utterances against the CPU oracle (C restatement of core.pyx, pinned by tests/test_oracle.py)."""
    from oracle import mas
    v = np.ascontiguousarray(values_t[idx].float().cpu().numpy())
    want = np.zeros(v.shape, np.int32)
    mas.maximum_path_c_port(want, v, np.ascontiguousarray(t_x[idx]), np.ascontiguousarray(t_y[idx]), omp=True)
    if field == "path":
        return bool(np.array_equal(result[idx].cpu().numpy().astype(np.int32), want))
    return bool(np.array_equal(result[idx].cpu().numpy(), want.sum(-1)))


def bench_mas_config(torch, dev, timer, ma, lib, name, b, tx, ty, t_x, t_y, peak, k=20):
    """One MAS configuration, device-resident, dense fp32 path out.  Returns the `configs` entry."""
    full = bool((t_x == tx).all() and (t_y == ty).all())
    cells = float((t_x.astype(np.int64) * t_y).sum())
    padded = float(b) * tx * ty
    algo = 4.0 * cells + 4.0 * padded                     # fp32 read of the in-band scores + dense fp32 path write
    per_set = int(padded) * 8
    nsets = rotating_sets(per_set) if per_set < (4 << 30) else 1          # one set of > 4 GB is already >> L2
    g = torch.Generator(device=dev).manual_seed(1234 + b + tx)
    vals = [torch.randn(b, tx, ty, generator=g, device=dev) for _ in range(nsets)]
    xl, yl = torch.from_numpy(t_x).to(dev), torch.from_numpy(t_y).to(dev)
    keep = [None] * nsets
    ones = torch.ones(1, 1, 1, device=dev).expand(b, tx, ty)             # full-length mask, never materialised

    def step(i):
        if full:
            keep[i % nsets] = ma.maximum_path(vals[i % nsets], ones)
        else:
            keep[i % nsets] = ma.maximum_path_lengths(vals[i % nsets], xl, yl)["path"]

    ms, launches, how = timer.run(step, k)
    out = ma.maximum_path_lengths(vals[0], xl, yl, dense=False, return_durations=True)
    rng = np.random.default_rng(7)
    idx = np.sort(rng.choice(b, min(b, 12), replace=False))
    ok = oracle_sample_ok(vals[0], t_x, t_y, out["durations"], idx)
    dense_ok = bool(torch.equal(keep[0].sum(-1).int(), out["durations"]))
    entry = {"name": name, "B": b, "T_text": tx, "T_mel": ty, "lengths": "full" if full else "mixed (t_x U{50..400}, t_y U{max(200,t_x)..2000}, seed 1239)",
             "ms": ms, "cells_per_s": cells / (ms * 1e-3), "utterances_per_sec": b / (ms * 1e-3),
             "roofline": {"bound": "hbm", "achieved": algo / (ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                          "frac": algo / (ms * 1e-3) / 1e9 / peak, "algorithmic_bytes_per_launch": algo},
             "kernel": lib.describe(b, tx, ty), "launches_timed": launches, "timing": how,
             "l2": ("rotating %d input/output sets (%.0f MB)" % (nsets, nsets * per_set / 1e6)) if nsets > 1 else "one set of %.1f GB >> L2" % (per_set / 1e9),
             "paths_equal_oracle": ok and dense_ok, "oracle_sample": int(len(idx))}
    del vals, keep
    torch.cuda.empty_cache()
    return entry


def bench_neg_cent(torch, dev, timer, ma, lib, peak_tflops):
    """Score-matrix kernels at the BASELINE shapes and the neg_cent -> MAS pipeline (BASELINE configs[1], [2])."""
    import aligner_b200.neg_cent as nc
    from oracle import neg_cent as nc_oracle
    out = {}
    g = torch.Generator(device=dev).manual_seed(1234 + 2)
    # ---- Gaussian prior, C2
    b, c, tx, ty = 64, 192, 200, 1000
    z = torch.randn(b, c, ty, generator=g, device=dev)
    m = torch.randn(b, c, tx, generator=g, device=dev)
    logs = torch.rand(b, c, tx, generator=g, device=dev) * 1.5 - 1.0           # U(-1, 0.5), SURVEY.md 8d
    ones = torch.ones(1, 1, 1, device=dev).expand(b, tx, ty)
    keep = [None, None, None, None]

    def nc_step(i):
        keep[i % 4] = nc.gaussian_neg_cent(z, m, logs)

    def pipe_step(i):
        keep[i % 4] = ma.maximum_path(nc.gaussian_neg_cent(z, m, logs), ones)

    ms, launches, _ = timer.run(nc_step, 20)
    score = nc.gaussian_neg_cent(z, m, logs)
    ref = nc_oracle.gaussian_neg_cent(z[:2].cpu().numpy(), m[:2].cpu().numpy(), logs[:2].cpu().numpy())
    err = float(np.abs(score[:2].cpu().numpy() - ref).max() / np.abs(ref).max())
    flop = 2.0 * 2 * c * b * tx * ty
    pms, _, _ = timer.run(pipe_step, 20)
    import aligner_b200.fused as fused
    xl = torch.full((b,), tx, dtype=torch.int32, device=dev)
    yl = torch.full((b,), ty, dtype=torch.int32, device=dev)

    def fused_step(i):
        keep[i % 4] = fused.gaussian_maximum_path(z, m, logs, x_lengths=xl, y_lengths=yl)

    fms, flaunches, _ = timer.run(fused_step, 20)
    fpath, fscore = fused.gaussian_maximum_path(z, m, logs, x_lengths=xl, y_lengths=yl)
    fused_ok = bool(torch.equal(fscore, score) and torch.equal(fpath, ma.maximum_path(score, ones)))
    out["gaussian_c2"] = {"shape": "B=%d C=%d T_text=%d T_mel=%d" % (b, c, tx, ty), "us": ms * 1e3, "launches_per_call": launches // 20 if launches else None,
                          "algorithmic_tflops": flop / (ms * 1e-3) / 1e12, "algorithmic_flop_per_cell": 4 * c,
                          "executed_tflops_3x_split": 3 * flop / (ms * 1e-3) / 1e12,
                          "max_rel_err_vs_fp64": err, "tolerance": 1e-5,
                          "tensor_pipe_source": "profiles/r02_nc_gauss.summary.txt (ncu --set full of this kernel; not measurable inside a timed run)",
                          "peak_bf16_tflops_measured": peak_tflops,
                          "pipeline_neg_cent_plus_mas_us": pms * 1e3,
                          "fused_entry_us": fms * 1e3, "fused_entry_launches": flaunches // 20 if flaunches else None,
                          "fused_entry_mode": "back to back, the search launched programmatically dependent on the score kernel (64 utterances need 64 of 148 SMs: pipelining measured no gain, DESIGN.md)",
                          "fused_bit_identical_to_separate_calls": fused_ok}
    # ---- the pipelined form of the fused entry pays when the search needs at most a third of the SMs: C3's shape with the Gaussian score
    b3, tx3, ty3 = 32, 300, 1500
    z3 = torch.randn(b3, c, ty3, generator=g, device=dev)
    m3 = torch.randn(b3, c, tx3, generator=g, device=dev)
    l3 = torch.rand(b3, c, tx3, generator=g, device=dev) * 1.5 - 1.0
    xl3 = torch.full((b3,), tx3, dtype=torch.int32, device=dev)
    yl3 = torch.full((b3,), ty3, dtype=torch.int32, device=dev)

    def sep3(i):
        keep[i % 4] = ma.maximum_path_lengths(nc.gaussian_neg_cent(z3, m3, l3), xl3, yl3)["path"]

    def fus3(i):
        keep[i % 4] = fused.gaussian_maximum_path(z3, m3, l3, x_lengths=xl3, y_lengths=yl3)

    s3, _, _ = timer.run(sep3, 20)
    f3, _, _ = timer.run(fus3, 20)
    p3, sc3 = fused.gaussian_maximum_path(z3, m3, l3, x_lengths=xl3, y_lengths=yl3)
    ok3 = bool(torch.equal(sc3, nc.gaussian_neg_cent(z3, m3, l3)) and torch.equal(p3, ma.maximum_path_lengths(sc3, xl3, yl3)["path"]))
    out["fused_gaussian_32x192x300x1500"] = {"separate_calls_us": s3 * 1e3, "fused_entry_us": f3 * 1e3,
                                             "fused_entry_mode": "pipelined: score kernel on 116 SMs publishes 128-frame tiles, the search runs beside it",
                                             "fused_bit_identical_to_separate_calls": ok3}
    del z, m, logs, score, z3, m3, l3
    # ---- OTA, C3
    b, c, tx, ty = 32, 80, 300, 1500
    q = torch.randn(b, c, ty, generator=g, device=dev)
    kk = torch.randn(b, c, tx, generator=g, device=dev)
    ones = torch.ones(1, 1, 1, device=dev).expand(b, tx, ty)

    def ota_step(i):
        keep[i % 4] = nc.ota_log_prob(q, kk)

    def ota_pipe(i):
        keep[i % 4] = ma.maximum_path(nc.ota_log_prob(q, kk), ones)

    ms, launches, _ = timer.run(ota_step, 20)
    score = nc.ota_log_prob(q, kk)
    ref = nc_oracle.ota_log_prob(q[:2].cpu().numpy(), kk[:2].cpu().numpy())
    err = float(np.abs(score[:2].cpu().numpy() - ref).max() / np.abs(ref).max())
    pms, _, _ = timer.run(ota_pipe, 20)
    out["ota_c3"] = {"shape": "B=%d C=%d T_text=%d T_mel=%d" % (b, c, tx, ty), "us": ms * 1e3, "launches_per_call": launches // 20 if launches else None,
                     "algorithmic_tflops": 2.0 * c * b * tx * ty / (ms * 1e-3) / 1e12, "out_write_GBps": 4.0 * b * tx * ty / (ms * 1e-3) / 1e9,
                     "max_rel_err_vs_fp64": err, "tolerance": 1e-5, "pipeline_neg_cent_plus_mas_us": pms * 1e3}
    del keep
    torch.cuda.empty_cache()
    return out


# --------------------------------------------------------------------------- strong scaling: one C5 batch split over the ranks
def c5_fill(torch, buf, ids, dev):
    """Scores of utterances `ids` (global indices) into buf[:len(ids)]: seeded per utterance, so that any rank can
    regenerate any utterance for the cross-check."""
    g = torch.Generator(device=dev)
    for j, i in enumerate(ids):
        g.manual_seed(977 * 1000003 + int(i))
        buf[j].normal_(generator=g)


def strong_scaling_c5(torch, dist, dev, ma, lib, rank, world, peak, total_b=8192, chunk=1024):
    from aligner_b200 import sharding
    t_x, t_y = c5_lengths(total_b)                         # every rank computes the same lengths and the same plan locally
    shards = sharding.balance_shards(t_x, t_y, world)
    loads = sharding.shard_loads(t_x, t_y, shards)
    mine = shards[rank]
    tx, ty = C5_TX, C5_TY
    chunk = min(chunk, max(1, len(mine)))
    buf = torch.empty(chunk, tx, ty, device=dev)           # 1024 x 400 x 2000 x 4 B = 3.3 GB of scores + 3.3 GB of path per chunk
    path = torch.empty(chunk, tx, ty, device=dev)
    dur_local = torch.zeros(len(mine), tx, dtype=torch.int32, device=dev)
    stream = torch.cuda.current_stream(dev).cuda_stream
    total_ms, nlaunch = 0.0, 0
    for c0 in range(0, len(mine), chunk):
        ids = mine[c0:c0 + chunk]
        n = len(ids)
        c5_fill(torch, buf, ids, dev)
        xl, yl = torch.from_numpy(t_x[ids]).to(dev), torch.from_numpy(t_y[ids]).to(dev)
        ws = ma._workspace(dev, stream, n, tx, ty)
        dur = torch.empty(n, tx, dtype=torch.int32, device=dev)

        def launch():
            lib.check(lib.lib.alb200_mas_device(buf.data_ptr(), xl.data_ptr(), yl.data_ptr(), path.data_ptr(), 4, 0x3F800000, 1, None, dur.data_ptr(),
                                                n, tx, ty, -1e9, ws.data_ptr(), ws.numel(), stream))
        launch()                                           # warm-up (same buffers: 6.6 GB per chunk >> L2)
        torch.cuda.synchronize()
        ts = []
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); launch(); e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
            nlaunch += 1
        total_ms += float(np.median(ts))
        dur_local[c0:c0 + n] = dur
    # ---- max over ranks of the device time
    t = torch.tensor([total_ms], device=dev, dtype=torch.float64)
    per_rank = [t.clone() for _ in range(world)]
    if world > 1:
        dist.all_gather(per_rank, t)
    per_rank_ms = [float(x.item()) for x in per_rank]
    ms_max = max(per_rank_ms)
    # ---- verification, outside the timed region: all-gather of the per-token durations (NCCL), neighbour sample re-run here
    nmax = max(len(s) for s in shards)
    padded = torch.zeros(nmax, tx, dtype=torch.int32, device=dev)
    padded[:len(mine)] = dur_local
    gathered = [torch.empty_like(padded) for _ in range(world)]
    if world > 1:
        dist.all_gather(gathered, padded)
    else:
        gathered = [padded]
    full = torch.zeros(total_b, tx, dtype=torch.int32, device=dev)
    for r, s in enumerate(shards):
        full[torch.from_numpy(s).to(dev)] = gathered[r][:len(s)]
    ok = bool((full.sum(1).cpu().numpy() == t_y).all())                         # every utterance: durations sum to its frame count
    nb = shards[(rank + 1) % world]
    rng = np.random.default_rng(5 + rank)
    ids = np.sort(rng.choice(nb, min(16, len(nb)), replace=False))
    c5_fill(torch, buf, ids, dev)
    mine_dur = ma.maximum_path_lengths(buf[:len(ids)], torch.from_numpy(t_x[ids]).to(dev), torch.from_numpy(t_y[ids]).to(dev),
                                       dense=False, return_durations=True)["durations"]
    ok = ok and bool(torch.equal(mine_dur, full[torch.from_numpy(ids).to(dev)]))
    oracle_ok = None
    if rank == 0:
        from oracle import mas
        v = np.ascontiguousarray(buf[:8].cpu().numpy())
        want = np.zeros(v.shape, np.int32)
        mas.maximum_path_c_port(want, v, np.ascontiguousarray(t_x[ids[:8]]), np.ascontiguousarray(t_y[ids[:8]]), omp=True)
        oracle_ok = bool(np.array_equal(mine_dur[:8].cpu().numpy(), want.sum(-1)))
    okt = torch.tensor([int(ok)], device=dev)
    if world > 1:
        dist.all_reduce(okt, op=dist.ReduceOp.MIN)
    del buf, path
    torch.cuda.empty_cache()
    cells = float((t_x.astype(np.int64) * t_y).sum())
    algo = 4.0 * cells + 4.0 * total_b * tx * ty
    return {"workload": "c5: ONE batch of B=%d mixed-length utterances (t_x U{50..400}, t_y U{max(200,t_x)..2000}, padded to 400x2000, seed 1239) "
                        "split across %d GPU(s) with balance_shards (cost t_x*t_y, longest first), chunks of <= %d utterances (6.6 GB) per launch" % (total_b, world, chunk),
            "scaling": "strong", "n_gpus": world, "ms": ms_max, "per_rank_ms": per_rank_ms,
            "cells_per_s": cells / (ms_max * 1e-3), "utterances_per_sec": total_b / (ms_max * 1e-3),
            "utterances_per_rank": [int(len(s)) for s in shards], "cells_per_rank": [int(x) for x in loads],
            "load_imbalance_max_over_mean": float(loads.max() / loads.mean()),
            "time_imbalance_max_over_mean": float(ms_max / np.mean(per_rank_ms)),
            "roofline_frac_aggregate": algo / (ms_max * 1e-3) / 1e9 / (peak * world),
            "launches_timed_per_rank": nlaunch, "collective_on_data_path": None,
            "durations_allgather_verified": bool(okt.item()), "oracle_sample_ok": oracle_ok}


def bind_to_gpu_numa_node(index: int):
    """Multi-rank runs: keep this rank's threads (and, by first touch, its pinned staging buffers) on the NUMA node of its GPU so
    that eight ranks' host-to-device streams do not all cross the same memory controller.  Returns the CPU list or None."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = [64 * w + bit for w, m in enumerate(mask) for bit in range(64) if (m >> bit) & 1]
        if cpus:
            os.sched_setaffinity(0, cpus)
            return cpus
    except Exception:
        pass
    return None


# --------------------------------------------------------------------------- ours
def run_ours(args, rank: int, world: int, local_rank: int):
    import torch
    import torch.distributed as dist

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa_cpus = bind_to_gpu_numa_node(local_rank) if world > 1 else None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    import aligner_b200.monotonic_align as ma
    from aligner_b200 import _lib
    from aligner_b200.monotonic_align.monotonic_align.core import maximum_path_c

    peak, peak_tflops, peak_src = peaks()
    timer = DeviceTimer(torch, dev, _lib, use_graph=not args.no_graph)
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
    clk = ClockSampler(local_rank)
    ms_step_local, launches, how = timer.run(step, args.steps, warmup=0, sampler=clk)
    barrier()
    t = torch.tensor([ms_step_local], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item())
    value = cells * world / (ms_step * 1e-3)

    # ---- the same launch timed eagerly with events around single launches (includes the launch gap; informational)
    kt = []
    for i in range(min(args.steps, 50)):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); step(i); e1.record()
        kt.append((e0, e1))
    torch.cuda.synchronize()
    kernel_ms_eager = float(np.median([a.elapsed_time(b_) for a, b_ in kt]))

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
    per_rank_e2e = [te.clone() for _ in range(world)]
    if world > 1:
        dist.all_gather(per_rank_e2e, te)
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
    kernel_desc = _lib.describe(b, tx, ty)
    del vals, keep
    torch.cuda.empty_cache()

    # ---- the other BASELINE configurations, the score kernels and the honest device-API comparison (one GPU)
    configs, neg, e2e_dev, cpu = None, None, None, None
    if world == 1 and not args.headline_only:
        configs = []
        for name in ("c1", "c3", "c4"):
            cb, ctx, cty, _ = WORKLOADS[name]
            configs.append(bench_mas_config(torch, dev, timer, ma, _lib, name, cb, ctx, cty, np.full(cb, ctx, np.int32), np.full(cb, cty, np.int32), peak))
        cx, cy = c5_lengths(2048)
        configs.append(bench_mas_config(torch, dev, timer, ma, _lib, "c5_B2048_mixed", 2048, C5_TX, C5_TY, cx, cy, peak, k=5))
        configs.append(bench_mas_config(torch, dev, timer, ma, _lib, "throughput_4096x200x1000", 4096, 200, 1000, np.full(4096, 200, np.int32),
                                        np.full(4096, 1000, np.int32), peak, k=10))
        try:
            neg = bench_neg_cent(torch, dev, timer, ma, _lib, peak_tflops)
        except Exception as exc:
            neg = {"error": repr(exc)}
    strong = None
    if not args.headline_only and not args.no_strong:
        barrier()
        strong = strong_scaling_c5(torch, dist, dev, ma, _lib, rank, world, peak, total_b=args.c5_batch)

    if rank == 0:
        achieved = BYTES_PER_CELL * cells / (ms_step_local * 1e-3) / 1e9
        traffic, traffic_src = None, None
        tj = ROOT / "profiles" / "traffic.json"
        if tj.exists():
            rec = json.loads(tj.read_text()).get(args.workload)
            if isinstance(rec, dict) and rec.get("kernel") == kernel_desc:      # only when the capture is of the very kernel shape launched here
                traffic, traffic_src = rec.get("dram_bytes"), rec.get("source")
        if world == 1 and not args.no_cpu:
            kind, fn, cores = load_cpu_reference("omp")
            reps = 8
            times, ref_paths = time_cpu(fn, values_np, t_x, t_y, reps, 2)
            parity = bool(np.array_equal(ref_paths, hp_np))
            cpu = {"value": cells / float(np.mean(times)), "unit": "cells/s", "cores": cores, "kind": kind,
                   "sample": "the whole %s batch, %d repetitions of maximum_path_c with -fopenmp on every host core (%.1f ms each)" % (args.workload, reps, 1e3 * float(np.mean(times))),
                   "best_ms": 1e3 * float(np.min(times)), "median_ms": 1e3 * float(np.median(times)), "paths_equal_gpu": parity}
            # (i) the reference as it ships: setup.py:5-9 passes no -fopenmp, the prange of core.pyx:44 is compiled out
            skind, sfn, _ = load_cpu_reference("serial")
            stimes, spaths = time_cpu(sfn, values_np, t_x, t_y, 4, 1)
            cpu["serial_as_shipped"] = {"value": cells / float(np.mean(stimes)), "unit": "cells/s", "cores": 1, "kind": skind,
                                        "best_ms": 1e3 * float(np.min(stimes)), "median_ms": 1e3 * float(np.median(stimes)),
                                        "paths_equal_gpu": bool(np.array_equal(spaths, hp_np))}
            # (ii) the reference's Python API fed CUDA tensors: its own 2 x D2H + H2D + staging (__init__.py:11-21) around its own core
            api = reference_api(fn)
            v_dev, m_dev = torch.from_numpy(values_np).to(dev), torch.ones(b, tx, ty, device=dev)
            rt = []
            for i in range(5):
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                rp = api(v_dev, m_dev)
                torch.cuda.synchronize()
                if i >= 1:
                    rt.append(time.perf_counter() - t0)
            cpu["reference_api_cuda_tensors"] = {"value": cells / float(np.mean(rt)), "unit": "cells/s", "cores": cores,
                                                 "kind": kind + " core + its __init__.py:11-21 staging restated line for line (the reference tree is not on the GPU box)",
                                                 "best_ms": 1e3 * float(np.min(rt)), "median_ms": 1e3 * float(np.median(rt))}
            # the same call through this repository's drop-in, wall clock with a synchronize per call (what a training step pays)
            dt_ = []
            for i in range(22):
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                op = ma.maximum_path(v_dev, m_dev)
                torch.cuda.synchronize()
                if i >= 2:
                    dt_.append(time.perf_counter() - t0)
            e2e_dev = {"value": cells / float(np.mean(dt_)), "unit": "cells/s", "best_ms": 1e3 * float(np.min(dt_)), "median_ms": 1e3 * float(np.median(dt_)),
                       "api": "monotonic_align.maximum_path(value, mask) on CUDA tensors, wall clock incl. launch + synchronize per call",
                       "paths_equal_reference_api": bool(torch.equal(op, rp)),
                       "speedup_vs_reference_api_cuda_tensors": float(np.mean(rt)) / float(np.mean(dt_))}
        line = {
            "metric": "mas_cells_per_sec", "value": value, "unit": "cells/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_string(args.workload),
                       "per_gpu": "every rank aligns its own batch of this shape (weak scaling); the strong-scaling leg is strong_scaling_c5",
                       "api": "monotonic_align.maximum_path(value, mask) on CUDA tensors, dense fp32 path out",
                       "l2": "rotating %d input/output sets (%.0f MB) > 3x L2, no flush" % (nsets, nsets * per_set / 1e6),
                       "launch": how, "kernel": kernel_desc},
            "utterances_per_sec": b * world / (ms_step * 1e-3),
            "clocks": clk.summary(),
            "e2e": {"value": e2e_val, "unit": "cells/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "api": "maximum_path_c(paths, values, t_xs, t_ys) with pinned host numpy buffers", "ms_per_step": float(te.item()) * 1e3,
                    "per_rank_ms": [float(x.item()) * 1e3 for x in per_rank_e2e],
                    "rank0_cpu_affinity": ("%d cpus of the GPU's NUMA node" % len(numa_cpus)) if numa_cpus else "unchanged",
                    "h2d_GBps_per_rank": [h2d / float(x.item()) / 1e9 for x in per_rank_e2e],
                    "paths_match_device_api": host_ok},
            "e2e_device_api": e2e_dev,
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                         "traffic_source": traffic_src, "kernel_ms": ms_step_local, "kernel_ms_eager_single_launch": kernel_ms_eager,
                         "peak_source": peak_src, "algorithmic_bytes_per_launch": BYTES_PER_CELL * cells},
            "cpu_baseline": cpu,
            "durations_allgather_verified": verified,
            "configs": configs,
            "neg_cent": neg,
            "strong_scaling_c5": strong,
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
    ap.add_argument("--headline-only", action="store_true", help="skip the configs / neg_cent / strong-scaling legs (used under ncu)")
    ap.add_argument("--no-strong", action="store_true", help="skip the C5 strong-scaling leg")
    ap.add_argument("--c5-batch", type=int, default=8192, help="utterances in the strong-scaling C5 batch")
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
