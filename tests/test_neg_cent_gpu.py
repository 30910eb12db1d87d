"""GPU: the score-matrix kernels against the fp64 oracle (oracle/neg_cent.py; parity unpinned, see its header),
and the north-star agreement criterion: MAS on the kernel's output vs MAS on the oracle's output."""
import numpy as np
import pytest
import torch

from oracle import mas as mas_oracle
from oracle import neg_cent as nc_oracle

pytestmark = pytest.mark.gpu

TOL = 1e-5      # north star: "within 1e-5 relative in fp32"; relative to the max |value| of the item, because
                # -0.5 z^2 s2 + z m s2 - 0.5 m^2 s2 cancels (neg_cent precision)


@pytest.fixture(scope="module")
def nc():
    import aligner_b200.neg_cent as m
    return m


def rel_err(got, want):
    finite = np.isfinite(want)
    assert (np.isfinite(got) == finite).all()
    scale = max(np.abs(want[finite]).max(), 1e-30) if finite.any() else 1.0
    if np.abs(want[finite]).max() == 0.0:
        return float(np.abs(got[finite]).max())
    return np.abs(got[finite] - want[finite]).max() / scale


@pytest.mark.parametrize("b,c,tx,ty", [(1, 1, 1, 1), (3, 7, 5, 9), (2, 192, 64, 128), (4, 192, 200, 1000), (2, 80, 130, 515), (1, 33, 301, 77)])
def test_gaussian_matches_fp64(nc, b, c, tx, ty):
    g = torch.Generator(device="cuda").manual_seed(1234 + tx)
    z = torch.randn(b, c, ty, generator=g, device="cuda")
    m = torch.randn(b, c, tx, generator=g, device="cuda")
    logs = torch.rand(b, c, tx, generator=g, device="cuda") * 1.5 - 1.0          # U(-1, 0.5), SURVEY.md 8d
    got = nc.gaussian_neg_cent(z, m, logs)
    assert got.shape == (b, tx, ty) and got.dtype == torch.float32
    want = nc_oracle.gaussian_neg_cent(z.cpu().numpy(), m.cpu().numpy(), logs.cpu().numpy())
    assert rel_err(got.cpu().numpy(), want) <= TOL
    again = nc.gaussian_neg_cent(z, m, logs)
    assert torch.equal(got, again), "not deterministic"


@pytest.mark.parametrize("b,c,tx,ty,with_prior,with_len", [(1, 1, 1, 1, False, False), (3, 80, 37, 150, True, True), (2, 80, 300, 1500, False, False),
                                                            (2, 80, 300, 333, True, False), (1, 16, 2000, 70, False, True)])
def test_ota_matches_fp64(nc, b, c, tx, ty, with_prior, with_len):
    g = torch.Generator(device="cuda").manual_seed(99 + tx)
    q = torch.randn(b, c, ty, generator=g, device="cuda")
    k = torch.randn(b, c, tx, generator=g, device="cuda")
    prior = None
    if with_prior:
        prior = torch.rand(b, tx, ty, generator=g, device="cuda")
    xl = None
    if with_len:
        xl = torch.randint(1, tx + 1, (b,), generator=g, device="cuda", dtype=torch.int32)
    got = nc.ota_log_prob(q, k, 0.0005, prior, xl)
    want = nc_oracle.ota_log_prob(q.cpu().numpy(), k.cpu().numpy(), 0.0005, None if prior is None else prior.cpu().numpy(),
                                  None if xl is None else xl.cpu().numpy())
    assert rel_err(got.cpu().numpy(), want) <= TOL
    if not with_prior:      # a log-softmax: every frame's column sums to one over the text axis
        p = torch.exp(got.double()).sum(1)
        assert torch.allclose(p, torch.ones_like(p), atol=1e-5)


@pytest.mark.parametrize("b,c,tx,ty", [(2, 192, 200, 1000), (2, 50, 600, 300), (1, 7, 513, 129)])
def test_tensor_core_and_cuda_core_paths_agree(nc, b, c, tx, ty):
    """The tcgen05 kernels (default) against the fixed-order FFMA kernels (ALB200_NC_FFMA=1): two independent implementations."""
    g = torch.Generator(device="cuda").manual_seed(7 + tx)
    z = torch.randn(b, c, ty, generator=g, device="cuda")
    m = torch.randn(b, c, tx, generator=g, device="cuda")
    logs = torch.rand(b, c, tx, generator=g, device="cuda") * 1.5 - 1.0
    tc_g, tc_o = nc.gaussian_neg_cent(z, m, logs), nc.ota_log_prob(z, m, 0.0005)
    from aligner_b200 import _lib
    _lib.set_option("nc_ffma", "1")
    try:
        ff_g, ff_o = nc.gaussian_neg_cent(z, m, logs), nc.ota_log_prob(z, m, 0.0005)
    finally:
        _lib.set_option("nc_ffma", None)
    assert (tc_g - ff_g).abs().max() <= TOL * ff_g.abs().max()
    assert (tc_o - ff_o).abs().max() <= TOL * ff_o.abs().max()


def test_three_generations_agree(nc):
    """fp16x3 / TMA (default), tf32x3 (option nc_v1) and FFMA (option nc_ffma): three independent implementations."""
    from aligner_b200 import _lib
    g = torch.Generator(device="cuda").manual_seed(17)
    b, c, tx, ty = 3, 192, 200, 1000
    z = torch.randn(b, c, ty, generator=g, device="cuda")
    m = torch.randn(b, c, tx, generator=g, device="cuda")
    logs = torch.rand(b, c, tx, generator=g, device="cuda") * 1.5 - 1.0
    q, k = torch.randn(b, 80, ty, generator=g, device="cuda"), torch.randn(b, 80, tx, generator=g, device="cuda")
    n0 = _lib.launch_count()
    g2, o2 = nc.gaussian_neg_cent(z, m, logs), nc.ota_log_prob(q, k)
    assert _lib.launch_count() - n0 == 4                     # prep + main kernel, twice: the TMA path really ran
    outs = {}
    for opt in ("nc_v1", "nc_ffma"):
        _lib.set_option(opt, "1")
        try:
            n0 = _lib.launch_count()
            outs[opt] = (nc.gaussian_neg_cent(z, m, logs), nc.ota_log_prob(q, k))
            assert _lib.launch_count() - n0 == 2
        finally:
            _lib.set_option(opt, None)
    for opt, (gg, oo) in outs.items():
        assert (g2 - gg).abs().max() <= TOL * gg.abs().max(), opt
        assert (o2 - oo).abs().max() <= TOL * oo.abs().max(), opt


@pytest.mark.parametrize("scale,expect_exact_path", [(1.0, False), (3000.0, True)])
def test_mel_side_outside_fp16_range_is_recomputed_exactly(nc, scale, expect_exact_path):
    """The mel side uses fixed power-of-two factors that are exact for |z| < 500 (|q| < 1000); a tile that sees more (or a
    non-finite value) is recomputed by its own CTA in plain fp32 -- slow, exact, never silent."""
    g = torch.Generator(device="cuda").manual_seed(23)
    b, c, tx, ty = 2, 64, 90, 520
    z = torch.randn(b, c, ty, generator=g, device="cuda")
    z[:, :, 130:140] *= scale                               # only the second mel tile of every utterance is affected
    m = torch.randn(b, c, tx, generator=g, device="cuda")
    logs = torch.rand(b, c, tx, generator=g, device="cuda") * 1.5 - 1.0
    got = nc.gaussian_neg_cent(z, m, logs).cpu().numpy()
    want = nc_oracle.gaussian_neg_cent(z.cpu().numpy(), m.cpu().numpy(), logs.cpu().numpy())
    for i in range(b):                                      # per mel tile: the huge frames must not hide errors elsewhere
        for y0 in range(0, ty, 128):
            assert rel_err(got[i, :, y0:y0 + 128], want[i, :, y0:y0 + 128]) <= TOL, (i, y0)
    q = torch.randn(b, 80, ty, generator=g, device="cuda")
    q[:, :, 300:310] *= scale
    k = torch.randn(b, 80, tx, generator=g, device="cuda")
    got = nc.ota_log_prob(q, k, 0.0005).cpu().numpy()
    if not expect_exact_path:
        want = nc_oracle.ota_log_prob(q.cpu().numpy(), k.cpu().numpy(), 0.0005)
    else:
        # |q| ~ 1e4: the sum of squared differences is ~1e10 and no fp32 evaluation reaches 1e-5 of fp64 any more; the exact path
        # must agree with the fixed-order fp32 FFMA kernel (same expression, same order) instead
        from aligner_b200 import _lib
        _lib.set_option("nc_ffma", "1")
        try:
            want = nc.ota_log_prob(q, k, 0.0005).cpu().numpy().astype(np.float64)
        finally:
            _lib.set_option("nc_ffma", None)
    for y0 in range(0, ty, 128):
        assert rel_err(got[:, :, y0:y0 + 128], want[:, :, y0:y0 + 128]) <= TOL, y0
    if expect_exact_path:                                   # NaN in, NaN out -- for the frames that hold it, not for the rest
        z[0, 3, 7] = float("nan")
        got = nc.gaussian_neg_cent(z, m, logs)
        assert torch.isnan(got[0, :, 7]).all() and torch.isfinite(got[0, :, 8]).all() and torch.isfinite(got[1]).all()


@pytest.mark.parametrize("b,c,tx,ty,scaling,native", [(3, 80, 60, 200, 1.0, True), (2, 80, 300, 1500, 1.0, True), (2, 40, 37, 132, 0.5, True),
                                                       (2, 16, 600, 640, 1.0, False),          # t_x > 512: prior materialised on the device
                                                       (2, 16, 50, 131, 2.0, False)])          # t_mel % 4 != 0: same
def test_ota_with_generated_beta_binomial_prior(nc, b, c, tx, ty, scaling, native):
    """the OTA prior BetaBinom(x; t_x - 1, s (y + 1), s (t_y - y)) generated inside the kernel against the
    fp64 oracle fed the dense scipy prior (oracle/neg_cent.py:beta_binomial_prior), ragged lengths."""
    from aligner_b200 import _lib
    g = torch.Generator(device="cuda").manual_seed(41 + tx)
    q = torch.randn(b, c, ty, generator=g, device="cuda")
    k = torch.randn(b, c, tx, generator=g, device="cuda")
    rng = np.random.default_rng(41 + tx)
    t_x = rng.integers(max(2, tx // 2), tx + 1, b).astype(np.int32)
    t_y = np.array([rng.integers(max(t_x[i], ty // 2), ty + 1) for i in range(b)], np.int32)
    t_x[0], t_y[0] = tx, ty
    dense = np.zeros((b, tx, ty))
    for i in range(b):
        dense[i, :t_x[i], :t_y[i]] = nc_oracle.beta_binomial_prior(int(t_x[i]), int(t_y[i]), scaling)
    want = nc_oracle.ota_log_prob(q.cpu().numpy(), k.cpu().numpy(), 0.0005, dense, t_x)
    n0 = _lib.launch_count()
    got = nc.ota_log_prob(q, k, 0.0005, x_lengths=torch.from_numpy(t_x).cuda(), y_lengths=torch.from_numpy(t_y).cuda(), prior_scaling=scaling)
    assert (_lib.launch_count() - n0 == 2) == native        # prep + score kernel, no other launch of ours: the prior was never materialised
    assert rel_err(got.cpu().numpy(), want) <= TOL
    # the dense prior built on the device for the other paths is the same function
    mine = nc.beta_binomial_prior(torch.from_numpy(t_x).cuda(), torch.from_numpy(t_y).cuda(), tx, ty, scaling).cpu().numpy()
    assert np.abs(mine - dense).max() <= 1e-6


def test_shape_fuzz_tma_path(nc):
    """Random small shapes through the TMA / tcgen05 kernels (t_mel % 4 == 0, t_x <= 512): channel tails, token tails around the
    16 / 256 boundaries, mel tiles shorter than 128 frames, one utterance to several rounds of tiles, ragged x_lengths."""
    from aligner_b200 import _lib
    rng = np.random.default_rng(777)
    g = torch.Generator(device="cuda").manual_seed(777)
    txs = [1, 2, 15, 16, 17, 31, 100, 255, 256, 257, 272, 300, 511, 512]
    for trial in range(40):
        b = int(rng.integers(1, 5))
        c = int(rng.choice([1, 3, 15, 16, 17, 31, 32, 33, 64, 80, 100, 192]))
        tx = int(rng.choice(txs))
        ty = 4 * int(rng.integers(1, 100))
        z = torch.randn(b, c, ty, generator=g, device="cuda")
        m = torch.randn(b, c, tx, generator=g, device="cuda")
        logs = torch.rand(b, c, tx, generator=g, device="cuda") * 1.5 - 1.0
        n0 = _lib.launch_count()
        got = nc.gaussian_neg_cent(z, m, logs)
        assert _lib.launch_count() - n0 == 2, (b, c, tx, ty)          # prep + TMA kernel
        want = nc_oracle.gaussian_neg_cent(z.cpu().numpy(), m.cpu().numpy(), logs.cpu().numpy())
        assert rel_err(got.cpu().numpy(), want) <= TOL, ("gaussian", b, c, tx, ty)
        xl = torch.from_numpy(rng.integers(1, tx + 1, b).astype(np.int32)).cuda()
        got = nc.ota_log_prob(z, m, 0.0005, None, xl)
        want = nc_oracle.ota_log_prob(z.cpu().numpy(), m.cpu().numpy(), 0.0005, None, xl.cpu().numpy())
        assert rel_err(got.cpu().numpy(), want) <= TOL, ("ota", b, c, tx, ty)


def test_c_entry_without_workspace(nc):
    """alb200_neg_cent_gaussian / _ota (no scratch argument) take the scratch from the stream-ordered allocator."""
    from aligner_b200 import _lib
    g = torch.Generator(device="cuda").manual_seed(29)
    b, c, tx, ty = 2, 40, 70, 260
    z = torch.randn(b, c, ty, generator=g, device="cuda")
    m = torch.randn(b, c, tx, generator=g, device="cuda")
    logs = torch.rand(b, c, tx, generator=g, device="cuda") - 0.5
    out = torch.empty(b, tx, ty, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    n0 = _lib.launch_count()
    _lib.check(_lib.lib.alb200_neg_cent_gaussian(z.data_ptr(), m.data_ptr(), logs.data_ptr(), out.data_ptr(), b, c, tx, ty, st))
    assert _lib.launch_count() - n0 == 2
    assert torch.equal(out, nc.gaussian_neg_cent(z, m, logs))
    _lib.check(_lib.lib.alb200_neg_cent_ota(z.data_ptr(), m.data_ptr(), None, None, out.data_ptr(), 0.0005, b, c, tx, ty, st))
    assert torch.equal(out, nc.ota_log_prob(z, m, 0.0005))


def _agreement(ma, score_gpu, score_ref64, t_x, t_y):
    ref32 = np.ascontiguousarray(score_ref64.astype(np.float32))
    want = np.zeros(ref32.shape, np.int32)
    mas_oracle.maximum_path_c_port(want, ref32.copy(), t_x, t_y, omp=True)
    got = ma.maximum_path_lengths(score_gpu, torch.from_numpy(t_x).cuda(), torch.from_numpy(t_y).cuda(), out_dtype=torch.int32)["path"]
    return float((got.cpu().numpy() == want).mean())


def test_paths_agree_gaussian(nc):
    """North star: paths from the fused score kernel agree with the reference formulation on >= 99.99% of cells."""
    import aligner_b200.monotonic_align as ma
    b, c, tx, ty = 16, 192, 200, 1000
    g = torch.Generator(device="cuda").manual_seed(5)
    z = torch.randn(b, c, ty, generator=g, device="cuda")
    m = torch.randn(b, c, tx, generator=g, device="cuda")
    logs = torch.rand(b, c, tx, generator=g, device="cuda") * 1.5 - 1.0
    rng = np.random.default_rng(5)
    t_x = rng.integers(50, tx + 1, b).astype(np.int32)
    t_y = np.array([rng.integers(max(200, t_x[i]), ty + 1) for i in range(b)], np.int32)
    score = nc.gaussian_neg_cent(z, m, logs)
    ref = nc_oracle.gaussian_neg_cent(z.cpu().numpy(), m.cpu().numpy(), logs.cpu().numpy())
    assert _agreement(ma, score, ref, t_x, t_y) >= 0.9999


def test_paths_agree_ota(nc):
    import aligner_b200.monotonic_align as ma
    b, c, tx, ty = 8, 80, 300, 1500
    g = torch.Generator(device="cuda").manual_seed(6)
    # keys/queries with real alignment structure: frames are noisy copies of the token they belong to
    k = torch.randn(b, c, tx, generator=g, device="cuda") * 3
    owner = torch.sort(torch.randint(0, tx, (b, ty), generator=g, device="cuda"), dim=1).values
    q = torch.gather(k, 2, owner[:, None, :].expand(b, c, ty)) + torch.randn(b, c, ty, generator=g, device="cuda")
    t_x, t_y = np.full(b, tx, np.int32), np.full(b, ty, np.int32)
    score = nc.ota_log_prob(q, k, 0.0005)
    ref = nc_oracle.ota_log_prob(q.cpu().numpy(), k.cpu().numpy(), 0.0005)
    assert _agreement(ma, score, ref, t_x, t_y) >= 0.9999


def test_rejects_cpu_tensors(nc):
    with pytest.raises(RuntimeError):
        nc.gaussian_neg_cent(torch.zeros(1, 2, 3), torch.zeros(1, 2, 4), torch.zeros(1, 2, 4))
