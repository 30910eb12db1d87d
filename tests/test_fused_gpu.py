"""GPU: fused score + search -- bit-identical to the two separate calls, and to the CPU oracle on its scores."""
import numpy as np
import pytest
import torch

from oracle import mas as mas_oracle
from oracle import neg_cent as nc_oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mods():
    import aligner_b200.fused as fused
    import aligner_b200.monotonic_align as ma
    import aligner_b200.neg_cent as nc
    from aligner_b200 import _lib
    return fused, ma, nc, _lib


def _inputs(seed, b, c, tx, ty):
    g = torch.Generator(device="cuda").manual_seed(seed)
    z = torch.randn(b, c, ty, generator=g, device="cuda")
    m = torch.randn(b, c, tx, generator=g, device="cuda")
    logs = torch.rand(b, c, tx, generator=g, device="cuda") * 1.5 - 1.0
    rng = np.random.default_rng(seed)
    t_x = rng.integers(max(1, tx // 4), tx + 1, b).astype(np.int32)
    t_y = np.array([rng.integers(max(t_x[i], ty // 4), ty + 1) for i in range(b)], np.int32)
    t_x[0], t_y[0] = tx, ty
    return z, m, logs, t_x, t_y


@pytest.mark.parametrize("mode", [None, "2"])          # default heuristic; pipelined whenever possible
@pytest.mark.parametrize("b,c,tx,ty", [(64, 192, 200, 1000),     # BASELINE configs[1]: the training-step shape
                                       (5, 80, 37, 132), (16, 192, 100, 800), (32, 64, 300, 1500), (3, 48, 500, 640),
                                       (2, 32, 700, 900),         # t_x > 512: the score kernel takes its fallback path
                                       (130, 64, 90, 400),        # too few SMs left over: back to back
                                       (400, 32, 60, 200),        # throughput regime: back to back
                                       (4, 40, 70, 261),          # t_mel % 4 != 0: fallback score kernel
                                       (4, 24, 96, 1600)])        # one compute warp, long mel axis: the search takes its 4-frame-lag form
def test_fused_equals_separate_calls(mods, b, c, tx, ty, mode):
    fused, ma, nc, _lib = mods
    _lib.set_option("fused_seq", mode)
    z, m, logs, t_x, t_y = _inputs(100 + b + tx, b, c, tx, ty)
    xl, yl = torch.from_numpy(t_x).cuda(), torch.from_numpy(t_y).cuda()
    score = nc.gaussian_neg_cent(z, m, logs)
    want = ma.maximum_path_lengths(score, xl, yl, return_durations=True)
    for rep in range(3):                                   # repeated: the tile flags and work counters re-arm themselves
        n0 = _lib.launch_count()
        path, neg_cent, dur = fused.gaussian_maximum_path(z, m, logs, x_lengths=xl, y_lengths=yl, return_durations=True)
        assert _lib.launch_count() - n0 == (3 if tx <= 512 and ty % 4 == 0 else 2)    # prep + score + search (fallback score kernel: one launch)
        assert torch.equal(neg_cent, score)
        assert torch.equal(path, want["path"]) and torch.equal(dur, want["durations"])
    # and against the CPU oracle run on the very same scores
    idx = np.arange(min(b, 6))
    ref = np.zeros((len(idx), tx, ty), np.int32)
    mas_oracle.maximum_path_c_port(ref, np.ascontiguousarray(score[:len(idx)].cpu().numpy()), t_x[idx].copy(), t_y[idx].copy(), omp=True)
    assert np.array_equal(path[:len(idx)].cpu().numpy().astype(np.int32), ref)
    _lib.set_option("fused_seq", None)


def test_fused_with_the_reference_style_mask(mods):
    fused, ma, nc, _lib = mods
    b, c, tx, ty = 6, 48, 90, 300
    z, m, logs, t_x, t_y = _inputs(7, b, c, tx, ty)
    mask = torch.zeros(b, tx, ty, device="cuda")
    for i in range(b):
        mask[i, :t_x[i], :t_y[i]] = 1
    path, neg_cent = fused.gaussian_maximum_path(z, m, logs, mask)
    assert torch.equal(path, ma.maximum_path(nc.gaussian_neg_cent(z, m, logs), mask))
    path_b, _ = fused.gaussian_maximum_path(z, m, logs, mask.bool())
    assert path_b.dtype == torch.float32 and torch.equal(path_b, path)
    # north star: >= 99.99 % of cells agree with the search run on the fp64 score
    ref64 = nc_oracle.gaussian_neg_cent(z.cpu().numpy(), m.cpu().numpy(), logs.cpu().numpy()).astype(np.float32)
    ref = np.zeros((b, tx, ty), np.int32)
    mas_oracle.maximum_path_c_port(ref, np.ascontiguousarray(ref64), t_x.copy(), t_y.copy(), omp=True)
    assert (path.cpu().numpy().astype(np.int32) == ref).mean() >= 0.9999


def test_fused_shape_fuzz(mods):
    """Random shapes and ragged lengths (empty items included) through the fused entry, pipelined whenever possible."""
    fused, ma, nc, _lib = mods
    rng = np.random.default_rng(4242)
    g = torch.Generator(device="cuda").manual_seed(4242)
    _lib.set_option("fused_seq", "2")
    try:
        for trial in range(30):
            b = int(rng.integers(1, 40))
            c = int(rng.choice([8, 33, 80, 192]))
            tx = int(rng.choice([1, 7, 31, 64, 130, 257, 300, 500]))
            ty = 4 * int(rng.integers(max(1, (tx + 3) // 4), max(2, (tx + 3) // 4) + 200))
            z = torch.randn(b, c, ty, generator=g, device="cuda")
            m = torch.randn(b, c, tx, generator=g, device="cuda")
            logs = torch.rand(b, c, tx, generator=g, device="cuda") * 1.5 - 1.0
            t_x = rng.integers(1, tx + 1, b).astype(np.int32)
            t_y = np.array([rng.integers(t_x[i], ty + 1) for i in range(b)], np.int32)
            if b > 2:
                t_x[1] = 0                                     # an empty item
            xl, yl = torch.from_numpy(t_x).cuda(), torch.from_numpy(t_y).cuda()
            score = nc.gaussian_neg_cent(z, m, logs)
            want = ma.maximum_path_lengths(score, xl, yl, return_durations=True)
            path, neg_cent, dur = fused.gaussian_maximum_path(z, m, logs, x_lengths=xl, y_lengths=yl, return_durations=True)
            assert torch.equal(neg_cent, score), (b, c, tx, ty)
            assert torch.equal(path, want["path"]) and torch.equal(dur, want["durations"]), (b, c, tx, ty)
    finally:
        _lib.set_option("fused_seq", None)


def test_fused_soak(mods):
    """Many back-to-back fused calls on alternating shapes: the publish / wait hand-off between the two concurrent kernels."""
    fused, ma, nc, _lib = mods
    _lib.set_option("fused_seq", "2")
    cases = []
    for seed, (b, c, tx, ty) in enumerate([(64, 192, 200, 1000), (20, 64, 120, 512), (7, 80, 260, 640)]):
        z, m, logs, t_x, t_y = _inputs(500 + seed, b, c, tx, ty)
        xl, yl = torch.from_numpy(t_x).cuda(), torch.from_numpy(t_y).cuda()
        want = ma.maximum_path_lengths(nc.gaussian_neg_cent(z, m, logs), xl, yl, dense=False, return_durations=True)["durations"]
        cases.append((z, m, logs, xl, yl, want))
    bad = torch.zeros((), dtype=torch.int64, device="cuda")
    for it in range(300):
        z, m, logs, xl, yl, want = cases[it % 3]
        _, _, dur = fused.gaussian_maximum_path(z, m, logs, x_lengths=xl, y_lengths=yl, return_durations=True)
        bad += (dur != want).any().long()
    _lib.set_option("fused_seq", None)
    assert int(bad.item()) == 0
