"""CPU: pins the oracle (oracle/mas_oracle.c) -- against the known answers of SURVEY.md 8c,
against golden vectors produced by the unmodified reference, and against the compiled
reference itself (oracle/_ref) on random inputs."""
import numpy as np
import pytest
import torch

from conftest import make_values, random_lengths
from oracle import mas


def _run_port(values, t_x, t_y, omp=False):
    p = np.zeros(values.shape, np.int32)
    v = values.copy()
    mas.maximum_path_c_port(p, v, t_x, t_y, omp=omp)
    return p, v


def test_known_answers():
    # (i) all ties, 4x8: diagonal then stay on the last token
    p, _ = _run_port(np.zeros((1, 4, 8), np.float32), np.array([4], np.int32), np.array([8], np.int32))
    want = np.zeros((4, 8), np.int32)
    want[0, 0] = want[1, 1] = want[2, 2] = 1
    want[3, 3:] = 1
    assert (p[0] == want).all()
    # (ii) square -> identity
    rng = np.random.default_rng(0)
    p, _ = _run_port(rng.standard_normal((1, 4, 4)).astype(np.float32), np.array([4], np.int32), np.array([4], np.int32))
    assert (p[0] == np.eye(4, dtype=np.int32)).all()
    # (iii) one token
    p, _ = _run_port(rng.standard_normal((1, 1, 6)).astype(np.float32), np.array([1], np.int32), np.array([6], np.int32))
    assert (p[0] == 1).all()
    # (iv) sentinel crossing: all cells -5e8, 3x6
    p, v = _run_port(np.full((1, 3, 6), -5e8, np.float32), np.array([3], np.int32), np.array([6], np.int32))
    want = np.zeros((3, 6), np.int32)
    want[0, :4] = 1
    want[1, 4] = 1
    want[2, 5] = 1
    assert (p[0] == want).all()
    assert v[0, 0, 0] == np.float32(-5e8) and v[0, 0, 1] == np.float32(-1e9) and v[0, 0, 2] == np.float32(-1.5e9)
    # (v) durations of the seeded 3x5x9 case
    torch.manual_seed(0)
    val = torch.randn(3, 5, 9).numpy()
    p, _ = _run_port(val, np.array([5, 3, 4], np.int32), np.array([9, 6, 4], np.int32))
    assert p.sum(-1).tolist() == [[1, 4, 2, 1, 1], [1, 1, 4, 0, 0], [1, 1, 1, 1, 0]]


def test_golden_vectors(golden):
    """Every golden case (outputs of the reference's own maximum_path) through the API-level restatement."""
    assert len(golden) >= 15
    for name, c in golden.items():
        value, mask = torch.from_numpy(c["value"]), torch.from_numpy(c["mask"])
        got = mas.maximum_path_port(value, mask)
        assert str(got.dtype) == str(c["path_dtype"]), name
        assert np.array_equal(got.numpy(), c["path"]), name


def test_bits_formulation_equals_table_form():
    rng = np.random.default_rng(5)
    for trial in range(120):
        b, tx = int(rng.integers(1, 4)), int(rng.integers(1, 70))
        ty = int(rng.integers(tx, 150))
        vals = make_values(rng, ["gauss", "ties", "sentinel", "negative"][trial % 4], (b, tx, ty))
        t_x, t_y = random_lengths(rng, b, tx, ty)
        p0, _ = _run_port(vals, t_x, t_y)
        p1, ftok = mas.mas_bits_port(vals, t_x, t_y)
        assert (p0 == p1).all()
        for i in range(b):
            assert (ftok[i, t_y[i]:] == -1).all()
            assert (p0[i].argmax(0)[: t_y[i]] == ftok[i, : t_y[i]]).all()


def test_nan_and_inf_follow_the_select_semantics():
    rng = np.random.default_rng(6)
    vals = rng.standard_normal((4, 12, 40)).astype(np.float32)
    vals[0, 3, 7] = np.nan
    vals[1, 2, 9] = -np.inf
    vals[2, 5, 20] = np.inf
    t_x, t_y = np.full(4, 12, np.int32), np.full(4, 40, np.int32)
    p0, _ = _run_port(vals, t_x, t_y)
    p1, _ = mas.mas_bits_port(vals, t_x, t_y)
    assert (p0 == p1).all()
    ref = mas.load_reference_core("serial")
    if ref is not None:
        p2 = np.zeros_like(p0)
        ref.maximum_path_c(p2, vals.copy(), t_x, t_y)
        assert (p0 == p2).all()


def test_openmp_build_identical():
    rng = np.random.default_rng(7)
    vals = rng.standard_normal((9, 33, 90)).astype(np.float32)
    t_x, t_y = random_lengths(rng, 9, 33, 90)
    assert (_run_port(vals, t_x, t_y)[0] == _run_port(vals, t_x, t_y, omp=True)[0]).all()


@pytest.mark.parametrize("kind", ["serial", "omp"])
def test_against_compiled_reference(kind):
    ref = mas.load_reference_core(kind)
    if ref is None:
        pytest.skip("oracle/_ref not built (needs /root/reference or a prebuilt copy)")
    rng = np.random.default_rng(8)
    for trial in range(150):
        b, tx = int(rng.integers(1, 5)), int(rng.integers(1, 60))
        ty = int(rng.integers(tx, 130))
        vals = make_values(rng, ["gauss", "ties", "sentinel", "negative", "zeros"][trial % 5], (b, tx, ty))
        t_x, t_y = random_lengths(rng, b, tx, ty)
        p_ref = np.zeros(vals.shape, np.int32)
        v_ref = vals.copy()
        ref.maximum_path_c(p_ref, v_ref, t_x, t_y)
        p, v = _run_port(vals, t_x, t_y)
        assert (p == p_ref).all()
        assert np.array_equal(v.view(np.int32), v_ref.view(np.int32))      # cumulative scores bit-identical too


def test_reference_python_api_when_available():
    api = mas.load_reference_api("serial")
    if api is None:
        pytest.skip("reference tree not present")
    g = torch.Generator().manual_seed(3)
    value = torch.randn(3, 10, 30, generator=g)
    mask = torch.zeros(3, 10, 30)
    for i, (a, b) in enumerate([(10, 30), (4, 17), (7, 7)]):
        mask[i, :a, :b] = 1
    assert torch.equal(api.maximum_path(value, mask), mas.maximum_path_port(value, mask))
