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


def test_sampled_parity_checker_agrees_with_the_full_oracle_and_detects_corruption():
    """oracle/sampled_parity.py (the checker bench.py and the full-size GPU tests use) cuts a sub-problem out of the
    graph for a sample of rows: on CPU tensors it must reproduce the full oracle's rows exactly and flag a wrong one."""
    import torch
    import sampled_parity as sp
    from synth_graphs import synth_graph
    n, k = 400, 16
    g = synth_graph(n, 9000, seed=5, kind="powerlaw")
    gen = torch.Generator().manual_seed(1)
    x, grad = torch.randn(n, 256, generator=gen), torch.rand(n, 256, generator=gen)
    deg = torch.clamp((g["indptr"][1:] - g["indptr"][:-1]).float(), min=1)
    vals, cols = oracle.topk(x.numpy(), k, 2)
    sel = cols.astype(np.uint8)
    ip, ix, va = g["indptr"].numpy(), g["indices"].numpy(), g["values"].numpy()
    out = torch.from_numpy(oracle.spgemm_fwd(ip, ix, va, vals, sel, deg=deg.numpy()))
    gs = torch.from_numpy(oracle.sspmm_bwd(ip, ix, va, grad.numpy(), sel, deg=deg.numpy()))
    tv, ts = torch.from_numpy(vals), torch.from_numpy(sel)
    rows = sp.sample_ids(n, 64, seed=3, device="cpu")
    assert sp.check_topk(x[rows], tv[rows], ts[rows], k) == (True, True)
    v, _ = sp.check_forward(g["indptr"], g["indices"], g["values"], tv, ts, rows, out[rows], row_div=deg)
    assert v == 0.0
    bad = out[rows].clone()
    bad[5, 17] += 1e-3
    assert sp.check_forward(g["indptr"], g["indices"], g["values"], tv, ts, rows, bad, row_div=deg)[0] > 1.0
    v, _, used = sp.check_backward(g["indptr"], g["indices"], g["values"], grad, ts, rows, gs[rows], row_div=deg)
    assert v == 0.0 and used == rows.numel()
    bad = gs[rows].clone()
    bad[7, 3] *= 1.001
    assert sp.check_backward(g["indptr"], g["indices"], g["values"], grad, ts, rows, bad, row_div=deg)[0] > 1.0
    wrong = ts.clone()
    wrong[int(rows[0]), 0] = (int(wrong[int(rows[0]), 0]) + 1) % 256          # a selector that is not in the top-k set
    assert sp.check_topk(x[rows], tv[rows], wrong[rows], k)[0] is False
    # a dense graph whose sample exceeds the edge budget is trimmed, not skipped
    v, _, used = sp.check_backward(g["indptr"], g["indices"], g["values"], grad, ts, rows, gs[rows], row_div=deg, max_edges=300)
    assert v == 0.0 and 0 < used < rows.numel()


def test_uint16_oracle_variants_reduce_to_the_uint8_ones():
    """oracle_spgemm_fwd16 / oracle_sspmm_bwd16 (feature widths above 256) are the same loops with a wider selector."""
    rng = np.random.default_rng(4)
    n, k = 120, 12
    deg = rng.integers(0, 9, n)
    indptr = np.zeros(n + 1, np.int32)
    indptr[1:] = np.cumsum(deg)
    indices = rng.integers(0, n, indptr[-1]).astype(np.int32)
    values = rng.random(indptr[-1]).astype(np.float32)
    x = rng.standard_normal((n, 256)).astype(np.float32)
    g = rng.random((n, 256)).astype(np.float32)
    vals, cols = oracle.topk(x, k, 1)
    assert np.array_equal(oracle.spgemm_fwd16(indptr, indices, values, vals, cols.astype(np.uint16), 256),
                          oracle.spgemm_fwd(indptr, indices, values, vals, cols.astype(np.uint8)))
    assert np.array_equal(oracle.sspmm_bwd16(indptr, indices, values, g, cols.astype(np.uint16)),
                          oracle.sspmm_bwd(indptr, indices, values, g, cols.astype(np.uint8)))
    # and a genuinely wide case against a dense numpy computation
    xw = rng.standard_normal((n, 600)).astype(np.float32)
    vw, cw = oracle.topk(xw, k, 1)
    dense = np.zeros((n, 600))
    np.put_along_axis(dense, cw.astype(np.int64), vw.astype(np.float64), 1)
    a = np.zeros((n, n))
    for r in range(n):
        for e in range(indptr[r], indptr[r + 1]):
            a[r, indices[e]] += values[e]
    np.testing.assert_allclose(oracle.spgemm_fwd16(indptr, indices, values, vw, cw.astype(np.uint16), 600), a @ dense, rtol=1e-6, atol=1e-6)
