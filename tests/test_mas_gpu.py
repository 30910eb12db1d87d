"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle, bit-exact.

Reference behaviour under test: monotonic_align/__init__.py:6-21 and core.pyx:7-45.
"""
import numpy as np
import pytest
import torch

from conftest import make_values, prefix_mask_np, random_lengths
from oracle import mas as oracle
from aligner_b200 import _lib

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ma():
    import aligner_b200.monotonic_align as m
    assert torch.cuda.is_available(), "GPU tests need a B200"
    return m


@pytest.fixture
def option():
    """Set library tuning options (alb200_set_option) for one test; defaults are restored afterwards."""
    touched = []

    def _set(name, value):
        _lib.set_option(name, value)
        touched.append(name)
    yield _set
    for name in touched:
        _lib.set_option(name, None)


def seed_of(*parts):
    """Deterministic across processes (unlike hash() of strings, which PYTHONHASHSEED randomises)."""
    import zlib
    return zlib.crc32(repr(parts).encode()) & 0x7fffffff


def reference_paths(values, t_x, t_y):
    """The reference's own compiled core.pyx (oracle/_ref, prebuilt, travels to the GPU box)."""
    ref = oracle.load_reference_core("omp") or oracle.load_reference_core("serial")
    if ref is None:
        pytest.skip("oracle/_ref is not built")
    p = np.zeros(values.shape, np.int32)
    ref.maximum_path_c(p, values.copy(), np.ascontiguousarray(t_x, np.int32), np.ascontiguousarray(t_y, np.int32))
    return p


def oracle_paths(values, t_x, t_y):
    p = np.zeros(values.shape, np.int32)
    oracle.maximum_path_c_port(p, values.copy(), np.ascontiguousarray(t_x, np.int32), np.ascontiguousarray(t_y, np.int32), omp=True)
    return p


def gpu_paths(ma, values, t_x, t_y, **kw):
    v = torch.from_numpy(values).cuda()
    out = ma.maximum_path_lengths(v, torch.from_numpy(t_x).cuda(), torch.from_numpy(t_y).cuda(), out_dtype=torch.int32,
                                  return_durations=True, return_frame_tokens=True, **kw)
    torch.cuda.synchronize()
    return out["path"].cpu().numpy(), out["durations"].cpu().numpy(), out["frame_tokens"].cpu().numpy()


def check_against_oracle(ma, values, t_x, t_y):
    want = oracle_paths(values, t_x, t_y)
    got, dur, ftok = gpu_paths(ma, values, t_x, t_y)
    bad = np.argwhere((got != want).reshape(len(t_x), -1).any(1)).ravel()
    assert bad.size == 0, "items differ: %s (t_x=%s t_y=%s)" % (bad[:8], t_x[bad[:8]], t_y[bad[:8]])
    assert (dur == want.sum(-1)).all()
    for i in range(len(t_x)):
        n = t_y[i] if (t_x[i] > 0 and t_y[i] >= t_x[i]) else 0
        assert (ftok[i, n:] == -1).all()
        if n:
            assert (ftok[i, :n] == want[i].argmax(0)[:n]).all()


# ------------------------------------------------------------------ golden + known answers
def test_golden_vectors_through_public_api(ma, golden):
    for name, c in golden.items():
        value, mask = torch.from_numpy(c["value"]).cuda(), torch.from_numpy(c["mask"]).cuda()
        v0 = value.clone()
        got = ma.maximum_path(value, mask)
        assert str(got.dtype) == str(c["path_dtype"]), name
        assert got.device == value.device and got.shape == value.shape
        assert not got.requires_grad
        assert torch.equal(value, v0), "input mutated: " + name
        assert np.array_equal(got.cpu().numpy(), c["path"]), name


# ------------------------------------------------------------------ differential, small shapes, every kernel shape
# "rows_per_lane,tile_frames,stages,bits_in_smem,skewed" -- every kernel shape in both forward forms
FORCES = [None, "1,32,2,1,0", "1,32,3,0,1", "2,16,2,1,0", "2,32,4,0,1", "2,32,2,1,1", "3,32,3,1,0", "3,32,3,0,1", "3,16,3,0,0", "4,32,2,1,1",
          "4,16,2,1,0", "4,32,2,1,0", "6,16,2,0,0", "6,32,2,1,1", "8,32,2,1,1", "8,16,2,0,0", "8,8,3,0,0", "16,16,2,0,0", "16,8,2,0,0"]


@pytest.mark.parametrize("force", FORCES)
@pytest.mark.parametrize("kind", ["gauss", "ties", "sentinel", "negative"])
def test_differential_small(ma, option, force, kind):
    if force:
        option("force", force)
    rmax = 4 * 32 * int(force.split(",")[0]) if force else 300         # a forced rows-per-lane caps t_x at 4 compute warps
    rng = np.random.default_rng(seed_of(force, kind))
    for trial in range(6):
        b = int(rng.integers(1, 9))
        tx = int(rng.integers(1, min(300, rmax)))
        ty = int(rng.integers(tx, 520))
        if trial % 2 == 0:
            ty = (ty + 3) // 4 * 4          # bulk-copy (aligned) path; odd trials take the unaligned loader
        values = make_values(rng, kind, (b, tx, ty))
        t_x, t_y = random_lengths(rng, b, tx, ty, full=(trial == 0))
        check_against_oracle(ma, values, t_x, t_y)


# The 4-frame-lag form (pre-skewed 3-D TMA boxes, forward_unit4): forced over every instance, one to four compute warps, a
# partly filled last warp, utterances shorter than the fill, ragged lengths; t_x must be a multiple of the rows per lane.
@pytest.mark.parametrize("rows,tx,ty", [(2, 64, 96), (2, 40, 400), (2, 200, 640), (2, 256, 300), (3, 96, 2000), (3, 150, 404), (3, 300, 700),
                                        (4, 128, 128), (4, 200, 1000), (4, 344, 520), (4, 512, 600), (2, 34, 36)])
@pytest.mark.parametrize("kind", ["gauss", "ties"])
def test_four_frame_lag_form(ma, option, rows, tx, ty, kind):
    option("force", "%d,32,0,-1,1,0,4" % rows)
    assert "form=skewed4" in _lib.describe(3, tx, ty)
    rng = np.random.default_rng(seed_of("lag4", rows, tx, ty, kind))
    for trial in range(2):
        b = int(rng.integers(1, 7))
        values = make_values(rng, kind, (b, tx, ty))
        t_x, t_y = random_lengths(rng, b, tx, ty, full=(trial == 0))
        check_against_oracle(ma, values, t_x, t_y)


@pytest.mark.parametrize("shape", [(6, 300, 640), (3, 500, 644), (20, 256, 400), (2, 400, 1500)])
def test_shared_zero_fill(ma, option, shape):
    """Fewer utterances than SMs and a big dense output: filler CTAs on the idle SMs zero the whole output from a shared cursor and the
    search CTAs scatter their ones only after every chunk is reported done.  Same paths as with every CTA filling its own
    utterance, for repeated launches (the counters in the workspace header re-arm themselves), empty utterances included."""
    b, tx, ty = shape
    rng = np.random.default_rng(seed_of("shared-zero", shape))
    values = make_values(rng, "gauss", (b, tx, ty))
    t_x, t_y = random_lengths(rng, b, tx, ty)
    if b > 2:
        t_x[1] = 0
    want = oracle_paths(values, t_x, t_y)
    v = torch.from_numpy(values).cuda()
    xl, yl = torch.from_numpy(t_x).cuda(), torch.from_numpy(t_y).cuda()
    for own in (None, "1", None):
        option("no_shared_zero", own)
        for _ in range(10):
            out = torch.full((b, tx, ty), 7.0, device="cuda")          # the kernel must overwrite every cell
            got = ma.maximum_path_lengths(v, xl, yl)["path"]
        assert np.array_equal(got.cpu().numpy(), want.astype(np.float32)), own


def test_two_tile_utterances_on_a_warm_device(ma):
    """An utterance whose compute warp needs only two tiles: both are requested before its lengths are known, so the compute warp
    can consume them and release their stages before the loader warp reaches its loop.  (The loader once waited on the stage's
    "empty" barrier there -- a phase that had already passed: a hang, but only on warm runs; found by the fused shape fuzz.)"""
    rng = np.random.default_rng(seed_of("two-tile"))
    for tx, ty in ((130, 704), (64, 96), (200, 1000)):
        values = make_values(rng, "gauss", (6, tx, ty))
        t_x = np.array([10, 1, 3, 31, 20, 5], np.int32)
        t_y = np.array([27, 1, 30, 32, 21, 33], np.int32)          # t_y - t_x + rows <= 32: one 32-frame word, two tiles
        want = oracle_paths(values, t_x, t_y)
        v = torch.from_numpy(values).cuda()
        xl, yl = torch.from_numpy(t_x).cuda(), torch.from_numpy(t_y).cuda()
        for _ in range(40):
            got = ma.maximum_path_lengths(v, xl, yl)["path"]
        assert np.array_equal(got.cpu().numpy(), want.astype(np.float32))


def test_four_frame_lag_is_chosen_for_one_long_warp(ma):
    """One compute warp and a long mel axis is where the form measured faster (mas_api.cu, choose_lag4)."""
    assert "form=skewed4" in _lib.describe(8, 96, 2000)
    assert "form=skewed4" not in _lib.describe(8, 200, 1000)
    assert "form=skewed4" not in _lib.describe(8, 95, 2000)          # 95 is not a multiple of 3 rows per lane
    rng = np.random.default_rng(seed_of("lag4-auto"))
    values = make_values(rng, "gauss", (5, 96, 2000))
    t_x, t_y = random_lengths(rng, 5, 96, 2000)
    check_against_oracle(ma, values, t_x, t_y)


@pytest.mark.parametrize("shape,native", [((6, 72, 190), True), ((3, 200, 403), True), ((2, 500, 640), True), ((9, 24, 33), True),
                                          ((6, 70, 190), False),        # t_text % 4 != 0: transposed on the device
                                          ((300, 72, 640), True),       # several SMs' worth: still the skewed form, persistent grid
                                          ((300, 72, 100), False),      # throughput regime
                                          ((2, 600, 700), False)])      # cluster shapes
def test_vits_layout_entry(ma, shape, native):
    """[b, t_mel, t_text] in and out, as VITS calls it; same search (its core indexes value[y, x])."""
    rng = np.random.default_rng(seed_of("vits", shape))
    b, tx, ty = shape
    values = make_values(rng, "gauss", (b, tx, ty))
    t_x, t_y = random_lengths(rng, b, tx, ty)
    want = oracle_paths(values, t_x, t_y)
    v = torch.from_numpy(np.ascontiguousarray(values.transpose(0, 2, 1))).cuda()                    # [b, t_mel, t_text]
    m = torch.from_numpy(np.ascontiguousarray(prefix_mask_np(t_x, t_y, tx, ty).transpose(0, 2, 1))).cuda()
    n0 = _lib.launch_count()
    got = ma.maximum_path_vits(v, m)
    assert got.shape == v.shape and got.dtype == v.dtype
    assert np.array_equal(got.cpu().numpy(), want.transpose(0, 2, 1).astype(np.float32))
    assert got.is_contiguous() == native                  # the native kernels write the [b, t_mel, t_text] result directly
    assert _lib.launch_count() == n0 + 1
    # bool mask, int result
    got = ma.maximum_path_vits(v, m.bool())
    assert np.array_equal(got.cpu().numpy(), want.transpose(0, 2, 1).astype(np.float32))


# ------------------------------------------------------------------ fp16 / bf16 scores read natively
@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("shape,native", [((8, 150, 400), True),      # latency regime, skewed/TMA form, 2 rows per lane
                                          ((5, 40, 96), True),        # one compute warp
                                          ((3, 500, 640), True),      # 4 rows per lane
                                          ((400, 90, 240), True),     # throughput regime, lock-step 16-frame tiles
                                          ((400, 90, 243), True),     # ... unaligned rows: element loader
                                          ((300, 600, 800), True),    # ... 8 rows per lane
                                          ((6, 150, 403), False),     # latency regime + unaligned rows: promoted on the device
                                          ((4, 700, 900), False)])    # latency regime, > 4 rows per lane: promoted on the device
def test_half_precision_scores(ma, dtype, shape, native):
    rng = np.random.default_rng(seed_of(str(dtype), shape))
    b, tx, ty = shape
    v = torch.from_numpy(make_values(rng, "gauss", shape)).cuda().to(dtype)
    t_x, t_y = random_lengths(rng, b, tx, ty)
    xl, yl = torch.from_numpy(t_x).cuda(), torch.from_numpy(t_y).cuda()
    want = oracle_paths(v.float().cpu().numpy(), t_x, t_y)          # the reference promotes with .astype(np.float32): exact
    out = ma.maximum_path_lengths(v, xl, yl, out_dtype=torch.int32, return_durations=True)
    assert np.array_equal(out["path"].cpu().numpy(), want)
    assert np.array_equal(out["durations"].cpu().numpy(), want.sum(-1))
    # the public API with a mask, result dtype = result_type(value, mask)
    mask = torch.from_numpy(prefix_mask_np(t_x, t_y, tx, ty)).cuda().to(dtype)
    got = ma.maximum_path(v, mask)
    assert got.dtype == dtype and np.array_equal(got.float().cpu().numpy(), want.astype(np.float32))
    # is the native kernel really taken (no promotion pass)?
    path = torch.empty(shape, dtype=torch.int32, device="cuda")
    stream = torch.cuda.current_stream().cuda_stream
    ws = ma._workspace(v.device, stream, b, tx, ty)
    rc = _lib.lib.alb200_mas_device_ex(v.data_ptr(), _lib.F16 if dtype == torch.float16 else _lib.BF16, xl.data_ptr(), yl.data_ptr(),
                                       None, 0, 0, 0, 0, path.data_ptr(), 4, 1, 1, None, None, None, b, tx, ty, -1e9,
                                       ws.data_ptr(), ws.numel(), stream)
    torch.cuda.synchronize()
    assert rc == (0 if native else _lib.E_UNSUPPORTED)
    if native:
        assert np.array_equal(path.cpu().numpy(), want)


def test_midsize_batch_runs_the_skewed_form_persistently(ma):
    """Several SMs' worth of utterances with t_x <= 512: one CTA per SM, skewed/TMA form, work cursor across many items."""
    rng = np.random.default_rng(91)
    for (b, tx, ty) in [(500, 150, 640), (420, 100, 600), (300, 400, 640), (300, 100, 400)]:
        assert "form=skewed" in _lib.describe(b, tx, ty)
        values = make_values(rng, "gauss", (b, tx, ty))
        t_x, t_y = random_lengths(rng, b, tx, ty)
        check_against_oracle(ma, values, t_x, t_y)
    assert "form=lockstep" in _lib.describe(4096, 100, 800)             # a full machine's worth of short utterances: occupancy-driven form
    assert "form=skewed" in _lib.describe(4096, 300, 1000)              # three rows per lane and a long mel axis: persistent skewed form at any batch size
    assert "form=lockstep" in _lib.describe(4096, 400, 1000)            # four rows per lane: not beyond five SMs' worth
    assert "form=skewed" in _lib.describe(500, 150, 400)                # 400 frames still amortise the pipeline fill (re-measured, mas_api.cu is_latency)
    assert "form=lockstep" in _lib.describe(500, 150, 360)              # shorter mel axis: they do not


@pytest.mark.parametrize("shape", [(1, 2048, 2304),      # cluster of 8 CTAs
                                   (1, 3000, 3200),      # 16 rows per lane, lock-step
                                   (2, 1, 5000), (1, 4, 20000),      # one token / very long mel axis
                                   (3, 513, 516),        # cluster of 3, almost square band
                                   (150, 513, 600),      # more clusters than fit: single-CTA throughput form
                                   (5, 33, 33), (2, 2047, 2050)])    # square; unaligned rows with 8 compute warps
def test_extreme_shapes(ma, shape):
    rng = np.random.default_rng(seed_of("extreme", shape))
    b, tx, ty = shape
    values = make_values(rng, "gauss", shape)
    t_x, t_y = random_lengths(rng, b, tx, ty)
    t_x[0], t_y[0] = tx, ty
    check_against_oracle(ma, values, t_x, t_y)


# ------------------------------------------------------------------ cluster mode: one utterance split over the CTAs of a cluster
@pytest.mark.parametrize("force,txmax", [("1,32,3,0,1,2", 256), ("2,32,3,0,1,2", 512), ("2,32,2,0,1,4", 1024), ("1,32,4,0,1,8", 1024),
                                         ("3,32,2,0,1,3", 1152)])
def test_cluster_mode_forced(ma, option, force, txmax):
    option("force", force)
    rng = np.random.default_rng(seed_of("cluster", force))
    for trial in range(3):
        b = int(rng.integers(1, 6))
        tx = int(rng.integers(txmax // 2, txmax + 1))
        ty = (int(rng.integers(tx, tx + 600)) + 3) // 4 * 4          # cluster mode needs 16-byte aligned rows
        values = make_values(rng, ["gauss", "ties", "sentinel"][trial], (b, tx, ty))
        t_x, t_y = random_lengths(rng, b, tx, ty, full=(trial == 0))
        assert "cluster=%s" % force.split(",")[5] in _lib.describe(b, tx, ty)
        check_against_oracle(ma, values, t_x, t_y)


def test_cluster_mode_is_the_default_for_long_text(ma):
    rng = np.random.default_rng(5)
    b, tx, ty = 3, 1000, 1400
    assert "cluster=4" in _lib.describe(b, tx, ty)
    values = make_values(rng, "gauss", (b, tx, ty))
    t_x, t_y = random_lengths(rng, b, tx, ty, full=True)
    check_against_oracle(ma, values, t_x, t_y)
    t_x, t_y = random_lengths(rng, b, tx, ty)                          # short items leave whole CTAs of a cluster idle
    t_x[0], t_y[0] = 7, 9
    check_against_oracle(ma, values, t_x, t_y)
    # the reference API (lengths from the mask, in every CTA of the cluster)
    want = oracle_paths(values, t_x, t_y)
    v = torch.from_numpy(values).cuda()
    got = ma.maximum_path(v, torch.from_numpy(prefix_mask_np(t_x, t_y, tx, ty)).cuda())
    assert np.array_equal(got.cpu().numpy(), want.astype(np.float32))
    assert "cluster=1" in _lib.describe(148, tx, ty)                   # not when the clusters would not all be resident


@pytest.mark.parametrize("skew", ["0", "1"])
def test_unaligned_loader_forced(ma, option, skew):
    option("force_unaligned", "1")
    option("force", "2,32,2,1," + skew)
    rng = np.random.default_rng(77)
    values = make_values(rng, "gauss", (5, 150, 400))
    t_x, t_y = random_lengths(rng, 5, 150, 400)
    check_against_oracle(ma, values, t_x, t_y)


@pytest.mark.parametrize("skew", ["0", "1"])
def test_both_forward_forms_on_ragged_batch(ma, option, skew):
    option("force", "2,32,3,1," + skew)
    rng = np.random.default_rng(78)
    values = make_values(rng, "ties", (40, 200, 600))
    t_x, t_y = random_lengths(rng, 40, 200, 600)
    check_against_oracle(ma, values, t_x, t_y)


# ------------------------------------------------------------------ BASELINE.json configs, full size
@pytest.mark.parametrize("b,tx,ty", [(16, 100, 800), (64, 200, 1000), (32, 300, 1500), (8, 1000, 6000)])
@pytest.mark.parametrize("kind", ["gauss", "ties"])
def test_baseline_configs_bit_exact(ma, b, tx, ty, kind):
    rng = np.random.default_rng(1234 + tx)
    values = make_values(rng, kind, (b, tx, ty))
    t_x, t_y = random_lengths(rng, b, tx, ty, full=True)
    check_against_oracle(ma, values, t_x, t_y)


def test_ragged_batch_work_stealing(ma):
    """More items than resident CTAs, mixed lengths (config 5 in miniature)."""
    rng = np.random.default_rng(99)
    b, tx, ty = 1500, 96, 256
    values = make_values(rng, "gauss", (b, tx, ty))
    t_x, t_y = random_lengths(rng, b, tx, ty)
    t_x[::50] = 0                       # sprinkle empty items
    check_against_oracle(ma, values, t_x, t_y)
    check_against_oracle(ma, values, t_x, t_y)   # second launch: the work counter re-armed itself


def test_config5_shape_properties(ma):
    """Mixed-length batch at the sweep's shape: invariants that need no oracle, plus a sampled oracle check."""
    g = torch.Generator(device="cuda").manual_seed(1239)
    b, tx, ty = 512, 400, 2000
    rng = np.random.default_rng(1239)
    t_x = rng.integers(50, tx + 1, b).astype(np.int32)
    t_y = np.array([rng.integers(max(200, t_x[i]), ty + 1) for i in range(b)], np.int32)
    values = torch.randn(b, tx, ty, generator=g, device="cuda")
    out = ma.maximum_path_lengths(values, torch.from_numpy(t_x).cuda(), torch.from_numpy(t_y).cuda(), out_dtype=torch.float32,
                                  return_durations=True, return_frame_tokens=True)
    path, dur, ftok = out["path"], out["durations"], out["frame_tokens"]
    txc, tyc = torch.from_numpy(t_x).cuda(), torch.from_numpy(t_y).cuda()
    assert ((path == 0) | (path == 1)).all()
    assert (path.sum(1).sum(1).int() == tyc).all()                       # exactly one token per real frame
    ys = torch.arange(ty, device="cuda")[None]
    assert (path.sum(1)[ys >= tyc[:, None]] == 0).all()                  # nothing past t_y
    assert (dur == path.sum(-1).int()).all() and (dur.sum(1) == tyc).all()
    assert (dur[torch.arange(tx, device="cuda")[None] >= txc[:, None]] == 0).all()
    inb = ys < tyc[:, None]
    assert (ftok[:, 0] == 0).all()
    assert (ftok.gather(1, (tyc - 1).long()[:, None])[:, 0] == txc - 1).all()   # ends on the last token
    step = ftok[:, 1:] - ftok[:, :-1]
    assert ((step == 0) | (step == 1))[inb[:, 1:]].all()                  # monotone, no skips
    idx = rng.choice(b, 24, replace=False)
    want = oracle_paths(values[idx].cpu().numpy(), t_x[idx], t_y[idx])
    assert np.array_equal(path[idx].cpu().numpy().astype(np.int32), want)


# ------------------------------------------------------------------ API behaviour
@pytest.mark.parametrize("vdt,mdt", [(torch.float32, torch.float32), (torch.float16, torch.float32), (torch.float64, torch.float32),
                                     (torch.float32, torch.bool), (torch.float32, torch.int64), (torch.float32, torch.float64),
                                     (torch.bfloat16, torch.bfloat16), (torch.float16, torch.float16), (torch.float32, torch.uint8)])
def test_api_dtypes(ma, vdt, mdt):
    rng = np.random.default_rng(31)
    b, tx, ty = 4, 45, 130
    t_x, t_y = random_lengths(rng, b, tx, ty)
    value = torch.from_numpy(make_values(rng, "gauss", (b, tx, ty))).to(vdt)
    mask = torch.from_numpy(prefix_mask_np(t_x, t_y, tx, ty)).to(mdt)
    got = ma.maximum_path(value.cuda(), mask.cuda())
    assert got.dtype == torch.result_type(value, mask)
    if vdt == torch.bfloat16:       # the reference raises TypeError at .numpy() for bf16; we accept it (exact promotion)
        want = oracle.maximum_path_port(value.float(), mask.float()).to(got.dtype)
    else:
        want = oracle.maximum_path_port(value, mask)
    assert torch.equal(got.cpu(), want)


def test_api_noncontiguous_and_expanded(ma):
    rng = np.random.default_rng(32)
    b, tx, ty = 3, 40, 96
    t_x, t_y = random_lengths(rng, b, tx, ty)
    base = torch.from_numpy(make_values(rng, "gauss", (b, ty, tx))).cuda()
    value = base.transpose(1, 2)                                         # non-contiguous view [b,tx,ty]
    xm = (torch.arange(tx)[None] < torch.from_numpy(t_x)[:, None]).float().cuda()
    ym = (torch.arange(ty)[None] < torch.from_numpy(t_y)[:, None]).float().cuda()
    mask = xm[:, :, None].expand(b, tx, ty) * ym[:, None, :].expand(b, tx, ty)
    got = ma.maximum_path(value, mask)
    want = oracle_paths(value.contiguous().cpu().numpy(), t_x, t_y)
    assert np.array_equal(got.cpu().numpy().astype(np.int32), want)
    # a stride-0 (expanded, never materialised) mask works too
    full = torch.ones(1, 1, 1, device="cuda").expand(b, tx, ty)
    got = ma.maximum_path(value, full)
    want = oracle_paths(value.contiguous().cpu().numpy(), np.full(b, tx, np.int32), np.full(b, ty, np.int32))
    assert np.array_equal(got.cpu().numpy().astype(np.int32), want)


def test_api_edge_shapes(ma):
    assert ma.maximum_path(torch.zeros(0, 5, 9).cuda(), torch.zeros(0, 5, 9).cuda()).shape == (0, 5, 9)
    v = torch.randn(2, 1, 1).cuda()
    assert torch.equal(ma.maximum_path(v, torch.ones_like(v)), torch.ones_like(v))
    with pytest.raises(RuntimeError):
        ma.maximum_path_lengths(torch.zeros(1, 2, 3), torch.tensor([2]), torch.tensor([3]))   # the extension entry is CUDA only
    with pytest.raises(ValueError):
        ma.maximum_path(torch.zeros(1, 2, 3).cuda(), torch.ones(1, 2, 4).cuda())


def test_invalid_lengths_flagged_not_crashing(ma):
    rng = np.random.default_rng(33)
    values = make_values(rng, "gauss", (3, 20, 30))
    t_x = np.array([20, 10, 5], np.int32)
    t_y = np.array([30, 4, 25], np.int32)                                 # item 1 has t_x > t_y
    ma.check_status()
    got, dur, ftok = gpu_paths(ma, values, t_x, t_y)
    assert ma.check_status() == 1 and ma.check_status() == 0
    assert got[1].sum() == 0 and (ftok[1] == -1).all()
    want = oracle_paths(values, np.array([20, 0, 5], np.int32), np.array([30, 0, 25], np.int32))
    assert np.array_equal(got, want)


def test_max_neg_val_keyword(ma):
    rng = np.random.default_rng(34)
    values = make_values(rng, "negative", (4, 30, 64)) * 10
    t_x, t_y = random_lengths(rng, 4, 30, 64)
    for neg in (-1e9, -100.0):
        want = np.zeros(values.shape, np.int32)
        oracle.maximum_path_c_port(want, values.copy(), t_x, t_y, max_neg_val=neg)
        got, _, _ = gpu_paths(ma, values, t_x, t_y, max_neg_val=neg)
        assert np.array_equal(got, want), neg


# ------------------------------------------------------------------ host-pointer entry = the reference's maximum_path_c
def test_host_maximum_path_c(ma):
    from monotonic_align.monotonic_align.core import maximum_path_c
    rng = np.random.default_rng(35)
    for (b, tx, ty) in [(1, 1, 1), (7, 50, 200), (64, 200, 1000), (40, 97, 333), (5, 700, 1204)]:      # the last one runs as clusters of 3 CTAs
        values = make_values(rng, "gauss", (b, tx, ty))
        t_x, t_y = random_lengths(rng, b, tx, ty)
        if b > 3:
            t_x[2] = 0
        v0 = values.copy()
        paths = np.zeros(values.shape, np.int32)
        assert maximum_path_c(paths, values, t_x, t_y) is None
        assert np.array_equal(values, v0)
        assert np.array_equal(paths, oracle_paths(values, t_x, t_y))
    with pytest.raises(ValueError):
        maximum_path_c(np.zeros((1, 2, 3), np.int64), np.zeros((1, 2, 3), np.float32), np.zeros(1, np.int32), np.zeros(1, np.int32))
    with pytest.raises(ValueError):
        maximum_path_c(np.zeros((1, 4, 3), np.int32), np.zeros((1, 4, 3), np.float32), np.array([4], np.int32), np.array([3], np.int32))
    # keyword form, like core.c:19803 allows
    paths = np.zeros((1, 2, 3), np.int32)
    maximum_path_c(paths=paths, values=np.zeros((1, 2, 3), np.float32), t_xs=np.array([2], np.int32), t_ys=np.array([3], np.int32), max_neg_val=-1e9)
    assert paths.sum() == 3


# ------------------------------------------------------------------ CPU tensors: the reference accepts any device (__init__.py:12-14,21)
def test_golden_vectors_on_cpu_tensors(ma, golden):
    """CPU tensors are staged through alb200_maximum_path_c (the search still runs on the B200) and come back on the CPU."""
    for name, c in golden.items():
        value, mask = torch.from_numpy(c["value"]), torch.from_numpy(c["mask"])
        v0 = value.clone()
        n0 = _lib.launch_count()
        got = ma.maximum_path(value, mask)
        assert got.device.type == "cpu" and str(got.dtype) == str(c["path_dtype"]) and got.shape == value.shape, name
        assert torch.equal(value, v0), "input mutated: " + name
        assert np.array_equal(got.numpy(), c["path"]), name
        if value.numel() and c["path"].any():
            assert _lib.launch_count() > n0, "no kernel launched for " + name


def test_apply_mask_reproduces_the_reference_for_arbitrary_masks(ma):
    """A mask that is not prefix-shaped: the reference multiplies first (__init__.py:11); apply_mask=True does the same."""
    rng = np.random.default_rng(36)
    b, tx, ty = 3, 24, 60
    value = torch.from_numpy(make_values(rng, "gauss", (b, tx, ty)))
    mask = torch.from_numpy(prefix_mask_np(np.array([24, 20, 11]), np.array([60, 44, 30]), tx, ty))
    mask[:, 3:6, 10:20] = 0                                   # holes inside the band
    want = oracle.maximum_path_port(value, mask)
    for dev in ("cuda", "cpu"):
        got = ma.maximum_path(value.to(dev), mask.to(dev), apply_mask=True)
        assert torch.equal(got.cpu(), want), dev


# ------------------------------------------------------------------ NaN / +-inf scores, against the reference's own compiled core
# Select semantics under test (core.c:19384-19391, 19444): `v_prev > v_cur ? v_prev : v_cur` -- a NaN v_prev is dropped, a NaN
# v_cur propagates; the backtrack's `<` is false on NaN.  OTA scores legitimately hold -inf rows (text padding).
NONFINITE_FORMS = [None, "2,32,3,1,0", "2,32,3,1,1", "1,32,3,0,1", "4,16,2,1,0", "4,32,2,1,1", "8,16,2,0,0", "2,32,3,0,1,2", "1,32,3,0,1,4"]


@pytest.mark.parametrize("force", NONFINITE_FORMS)
@pytest.mark.parametrize("kind", ["nan", "pinf", "ninf", "ota_rows", "mixed"])
def test_nonfinite_scores_match_the_compiled_reference(ma, option, force, kind):
    if force:
        option("force", force)
    parts = [int(q) for q in force.split(",")] if force else None
    nc = parts[5] if parts and len(parts) > 5 else 1
    rmax = 4 * 32 * parts[0] * nc if parts else 300
    rng = np.random.default_rng(seed_of("nonfinite", force, kind))
    for trial in range(4):
        b = int(rng.integers(1, 7))
        tx = int(rng.integers(max(2, rmax // 2 if nc > 1 else 2), min(300 if nc == 1 else rmax, rmax) + 1))
        ty = (int(rng.integers(tx, tx + 400)) + 3) // 4 * 4
        values = make_values(rng, "gauss", (b, tx, ty))
        t_x, t_y = random_lengths(rng, b, tx, ty, full=(trial == 0))
        n = max(1, values.size // 200)
        flat = values.reshape(-1)
        idx = rng.choice(values.size, n, replace=False)
        if kind == "nan":
            flat[idx] = np.nan
        elif kind == "pinf":
            flat[idx] = np.inf
        elif kind == "ninf":
            flat[idx] = -np.inf
        elif kind == "ota_rows":                       # whole token rows at -inf, like OTA text padding inside the tensor
            for i in range(b):
                values[i, rng.integers(0, tx, max(1, tx // 10))] = -np.inf
        else:
            flat[idx] = rng.choice(np.array([np.nan, np.inf, -np.inf], np.float32), n)
        want = reference_paths(values, t_x, t_y)
        got, dur, _ = gpu_paths(ma, values, t_x, t_y)
        bad = np.argwhere((got != want).reshape(b, -1).any(1)).ravel()
        assert bad.size == 0, "items differ from the compiled reference: %s (t_x=%s t_y=%s)" % (bad[:8], t_x[bad[:8]], t_y[bad[:8]])
        assert (dur == want.sum(-1)).all()
        assert np.array_equal(want, oracle_paths(values, t_x, t_y))     # and the C restatement agrees with the reference here too


def test_compiled_reference_directly_on_baseline_shape(ma):
    """GPU path vs oracle/_ref itself (not the restatement) at the bench shape, ragged."""
    rng = np.random.default_rng(1234 + 2)
    b, tx, ty = 64, 200, 1000
    values = make_values(rng, "gauss", (b, tx, ty))
    t_x, t_y = random_lengths(rng, b, tx, ty)
    t_x[:8], t_y[:8] = tx, ty
    want = reference_paths(values, t_x, t_y)
    got, dur, _ = gpu_paths(ma, values, t_x, t_y)
    assert np.array_equal(got, want) and (dur == want.sum(-1)).all()


# ------------------------------------------------------------------ config 5 through the shard planner, on one GPU
def test_c5_ragged_batch_through_balance_shards(ma):
    """A C5-shaped batch is split with balance_shards into 1/2/4/8 shards, the shards run one after the other on this GPU,
    and the reassembled durations / frame tokens equal the unsharded run (utterance independence, core.pyx:44-45)."""
    from aligner_b200 import sharding
    rng = np.random.default_rng(1239)
    b, tx, ty = 384, 400, 2000
    t_x = rng.integers(50, tx + 1, b).astype(np.int32)
    t_y = np.array([rng.integers(max(200, t_x[i]), ty + 1) for i in range(b)], np.int32)
    g = torch.Generator(device="cuda").manual_seed(1239)
    values = torch.randn(b, tx, ty, generator=g, device="cuda")
    xl, yl = torch.from_numpy(t_x).cuda(), torch.from_numpy(t_y).cuda()
    whole = ma.maximum_path_lengths(values, xl, yl, dense=False, return_durations=True, return_frame_tokens=True)
    for world in (2, 4, 8):
        shards = sharding.balance_shards(t_x, t_y, world)
        loads = sharding.shard_loads(t_x, t_y, shards)
        assert loads.max() - loads.min() <= sharding.item_cost(t_x, t_y).max()
        dur = torch.full_like(whole["durations"], -7)
        ftok = torch.full_like(whole["frame_tokens"], -7)
        for s in shards:
            idx = torch.from_numpy(s).cuda()
            out = ma.maximum_path_lengths(values[idx], xl[idx], yl[idx], dense=False, return_durations=True, return_frame_tokens=True)
            dur[idx] = out["durations"]
            ftok[idx] = out["frame_tokens"]
        assert torch.equal(dur, whole["durations"]) and torch.equal(ftok, whole["frame_tokens"]), world
    sample = rng.choice(b, 12, replace=False)
    want = oracle_paths(values[sample].cpu().numpy(), t_x[sample], t_y[sample])
    assert np.array_equal(whole["durations"][sample].cpu().numpy(), want.sum(-1))


# ------------------------------------------------------------------ soak: the flag hand-offs between warps, 10^4 launches
def test_soak_ten_thousand_launches(ma):
    """Random small shapes in every regime, 10^4 launches: every result is compared with the C restatement (durations and
    frame tokens) -- a lost or reordered hand-off between warps (boundary ring, walker -> emitters) would show up here."""
    rng = np.random.default_rng(20261017)
    pool = []
    for _ in range(40):
        b = int(rng.choice([1, 3, 8, 40, 180, 400]))
        tx = int(rng.integers(1, 260))
        ty = int(rng.integers(tx, tx + 300))
        if rng.integers(0, 2):
            ty = (ty + 3) // 4 * 4
        values = make_values(rng, ["gauss", "ties"][int(rng.integers(0, 2))], (b, tx, ty))
        t_x, t_y = random_lengths(rng, b, tx, ty)
        _, ftok = oracle.mas_bits_port(values, t_x, t_y, omp=True)
        for i in range(b):
            ftok[i, t_y[i]:] = -1
        pool.append((torch.from_numpy(values).cuda(), torch.from_numpy(t_x).cuda(), torch.from_numpy(t_y).cuda(), torch.from_numpy(ftok).cuda()))
    bad = torch.zeros((), dtype=torch.int64, device="cuda")
    n = 10000
    for it in range(n):
        v, xl, yl, want = pool[int(rng.integers(0, len(pool)))]
        out = ma.maximum_path_lengths(v, xl, yl, dense=False, return_frame_tokens=True)
        bad += (out["frame_tokens"] != want).any().long()
    assert int(bad.item()) == 0, "%d of %d launches differed from the oracle" % (int(bad.item()), n)
