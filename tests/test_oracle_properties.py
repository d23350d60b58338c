"""CPU: hypothesis property tests of the oracle (random CSR / CBSR, ties, empty and ragged rows)."""
import numpy as np
from hypothesis import given, settings, strategies as st

import oracle


@st.composite
def csr_problem(draw):
    n = draw(st.integers(1, 40))
    k = draw(st.sampled_from([1, 3, 8, 16, 19, 32]))
    seed = draw(st.integers(0, 2 ** 31 - 1))
    rng = np.random.default_rng(seed)
    deg = rng.integers(0, draw(st.integers(1, 90)), n)
    deg[rng.random(n) < 0.2] = 0                                   # empty rows
    indptr = np.zeros(n + 1, np.int32)
    indptr[1:] = np.cumsum(deg)
    indices = rng.integers(0, n, indptr[-1]).astype(np.int32)
    values = rng.standard_normal(indptr[-1]).astype(np.float32)
    levels = draw(st.sampled_from([0, 3, 17]))                     # 0: continuous values, else heavy ties
    x = rng.standard_normal((n, 256)).astype(np.float32)
    if levels:
        x = np.round(x * levels) / levels
    g = rng.standard_normal((n, 256)).astype(np.float32)
    return indptr, indices, values, x, g, k


@settings(max_examples=40, deadline=None)
@given(csr_problem())
def test_topk_definition(p):
    _, _, _, x, _, k = p
    v0, c0 = oracle.topk(x, k, 0)
    for r in range(x.shape[0]):
        order = sorted(range(256), key=lambda j: (-x[r, j] if x[r, j] != 0 else 0.0, j))   # value desc (-0 == 0), column asc
        assert c0[r].tolist() == order[:k]
    for o in (1, 2):
        v, c = oracle.topk(x, k, o)
        assert np.array_equal(np.sort(c, 1), np.sort(c0, 1))
        assert np.array_equal(np.take_along_axis(x, c.astype(np.int64), 1), v)


@settings(max_examples=40, deadline=None)
@given(csr_problem())
def test_forward_backward_against_scipy_and_adjoint(p):
    indptr, indices, values, x, g, k = p
    vals, cols = oracle.topk(x, k, 2)
    sel = cols.astype(np.uint8)
    out = oracle.spgemm_fwd(indptr, indices, values, vals, sel)
    gs = oracle.sspmm_bwd(indptr, indices, values, g, sel)
    np.testing.assert_allclose(out, oracle.spgemm_fwd_scipy(indptr, indices, values, vals, cols), rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(gs, oracle.sspmm_bwd_scipy(indptr, indices, values, g, cols), rtol=1e-5, atol=1e-5)
    empty = np.diff(indptr) == 0
    assert not out[empty].any()                                     # rows without edges stay exactly zero
    lhs = float((out.astype(np.float64) * g).sum())
    rhs = float((vals.astype(np.float64) * gs).sum())
    assert abs(lhs - rhs) <= 1e-3 * (1.0 + abs(lhs))
    quads, w = oracle.warp4(indptr, 64)
    out4 = oracle.spgemm_fwd_warp4(quads, indices, values, vals, sel, len(indptr) - 1)
    np.testing.assert_allclose(out4, out, rtol=1e-6, atol=1e-6)
    q = quads.reshape(-1, 4)
    assert w == int(np.ceil(np.diff(indptr) / 64).sum()) and (q[:, 2] >= 1).all() and (q[:, 2] <= 64).all()
