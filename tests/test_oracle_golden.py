"""CPU: pins the oracle (oracle/maxk_oracle.c) against outputs of the REFERENCE itself.

ref_py.npz   -- the reference's Python code run on the CPU  (tests/golden/make_golden_py.py)
ref_cuda.npz -- the reference's CUDA kernels, unmodified, run on a B200 (tests/golden/make_golden_cuda.py)
The reference ships no golden vectors of its own for this path.
"""
import os

import numpy as np
import pytest

import oracle
from helpers import assert_close

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def ref_py():
    return np.load(os.path.join(GOLD, "ref_py.npz"))


@pytest.fixture(scope="module")
def ref_cuda():
    return np.load(os.path.join(GOLD, "ref_cuda.npz"))


def test_topk_then_forward_matches_reference_python_operator(ref_py):
    """maxk_spgemm_function.py forward (torch.topk -> scatter_ -> sparse.mm -> /in_degrees)."""
    k = int(ref_py["k"])
    for order in (0, 1, 2):       # the contraction does not depend on the order of a row's entries
        vals, cols = oracle.topk(ref_py["x"], k, order)
        sel = cols.astype(np.uint8)
        raw = oracle.spgemm_fwd(ref_py["indptr"], ref_py["indices"], ref_py["values"], vals, sel)
        assert_close(raw, ref_py["spgemm_fwd_raw"], "forward raw (order %d)" % order)
        norm = oracle.spgemm_fwd(ref_py["indptr"], ref_py["indices"], ref_py["values"], vals, sel, deg=ref_py["in_deg"])
        assert_close(norm, ref_py["spgemm_fwd_norm"], "forward / in_degrees (order %d)" % order)


@pytest.mark.parametrize("k", [8, 32])
def test_maxk_activation_matches_reference_classes(ref_py, k):
    x, up = ref_py["maxk_x"], ref_py["maxk_up"]
    vals, cols = oracle.topk(x, k, 0)
    assert np.array_equal(vals, ref_py["optmaxk_vals_k%d" % k])             # torch.topk values, sorted desc
    assert np.array_equal(cols.astype(np.int64), ref_py["optmaxk_idx_k%d" % k])
    assert np.array_equal(oracle.maxk_act_fwd(x, cols), ref_py["maxk_fwd_k%d" % k])
    assert np.array_equal(oracle.maxk_act_fwd(x, cols), ref_py["optmaxk_fwd_k%d" % k])
    assert np.array_equal(oracle.maxk_act_bwd(up, cols), ref_py["maxk_bwd_k%d" % k])
    # OPTMaxK.backward in the reference ignores grad_topk_values (model_integrated_v3.py:40-43)
    assert np.array_equal(oracle.maxk_act_bwd(up, cols), ref_py["optmaxk_bwd_k%d" % k])


def test_warp4_matches_reference_generate_meta(ref_py):
    got, w = oracle.warp4(ref_py["indptr"], 64)
    assert np.array_equal(got, ref_py["warp4_small"]) and w == ref_py["warp4_small"].size // 4
    got, w = oracle.warp4(ref_py["big_indptr"], 64)
    assert np.array_equal(got, ref_py["warp4_big"])
    assert (got.reshape(-1, 4)[:, 2] <= 64).all() and got.reshape(-1, 4)[:, 2].max() == 64


@pytest.mark.parametrize("name", ["k32", "k64"])
def test_oracle_matches_reference_cuda_kernels(ref_cuda, name):
    g = {key[len(name) + 1:]: ref_cuda[key] for key in ref_cuda.files if key.startswith(name + "_")}
    n = g["indptr"].size - 1
    fwd = oracle.spgemm_fwd(g["indptr"], g["indices"], g["values"], g["data"], g["sel"])
    assert_close(fwd, g["fwd"], "oracle fwd vs spmm_kernel_opt2_sparse_v3")
    fwd4 = oracle.spgemm_fwd_warp4(g["warp4"], g["indices"], g["values"], g["data"], g["sel"], n)
    assert_close(fwd4, g["fwd"], "oracle fwd (warp4-driven) vs spmm_kernel_opt2_sparse_v3")
    bwd = oracle.sspmm_bwd(g["indptr"], g["indices"], g["values"], g["grad"], g["sel"])
    assert_close(bwd, g["bwd"], "oracle bwd vs spmm_kernel_opt2_sparse_backward_v3")


def test_c_oracle_agrees_with_independent_scipy_restatement():
    rng = np.random.default_rng(5)
    n, k = 400, 16
    deg = rng.integers(0, 60, n)
    indptr = np.zeros(n + 1, np.int32)
    indptr[1:] = np.cumsum(deg)
    indices = rng.integers(0, n, indptr[-1]).astype(np.int32)
    values = rng.standard_normal(indptr[-1]).astype(np.float32)
    x = rng.standard_normal((n, 256)).astype(np.float32)
    g = rng.standard_normal((n, 256)).astype(np.float32)
    vals, cols = oracle.topk(x, k, 2)
    sel = cols.astype(np.uint8)
    assert_close(oracle.spgemm_fwd(indptr, indices, values, vals, sel),
                 oracle.spgemm_fwd_scipy(indptr, indices, values, vals, cols), "fwd C vs scipy", rtol=1e-6)
    assert_close(oracle.sspmm_bwd(indptr, indices, values, g, sel),
                 oracle.sspmm_bwd_scipy(indptr, indices, values, g, cols), "bwd C vs scipy", rtol=1e-6)


def test_topk_orders_are_permutations_and_ties_take_lowest_column():
    x = np.array([[1, 3, 3, 3, 2, 3, 0, 3] * 32], np.float32)        # SURVEY.md hard part 3 probe
    v0, c0 = oracle.topk(x, 2, 0)
    assert c0.tolist() == [[1, 2]]                                     # torch.topk gives [1, 7] here
    rng = np.random.default_rng(1)
    y = rng.standard_normal((50, 256)).astype(np.float32)
    for k in (8, 16, 32, 64, 19):
        a, b, c = (oracle.topk(y, k, o) for o in (0, 1, 2))
        assert np.array_equal(np.sort(a[1], 1), b[1]) and np.array_equal(np.sort(c[1], 1), b[1])
        if oracle.banked_modulus(k) == 1:                              # no vectorised path: plain column order
            assert np.array_equal(c[1], b[1])
        elif k < 32:                                                   # classes mod 4, largest first, columns ascending
            for row in c[1]:
                size = np.bincount(row % 4, minlength=4)
                cls = row % 4
                assert all(size[cls[i]] > size[cls[i + 1]] or (size[cls[i]] == size[cls[i + 1]] and cls[i] < cls[i + 1])
                           or (cls[i] == cls[i + 1] and row[i] < row[i + 1]) for i in range(k - 1))
        assert np.array_equal(np.take_along_axis(y, c[1].astype(np.int64), 1), c[0])


def test_forward_backward_are_adjoint():
    """<A x_s, g> == <x_vals, sample(A^T g)>: the property the autograd pair relies on."""
    rng = np.random.default_rng(2)
    n, k = 300, 32
    deg = rng.integers(0, 40, n)
    indptr = np.zeros(n + 1, np.int32)
    indptr[1:] = np.cumsum(deg)
    indices = rng.integers(0, n, indptr[-1]).astype(np.int32)
    values = rng.random(indptr[-1]).astype(np.float32)
    x = rng.standard_normal((n, 256)).astype(np.float32)
    g = rng.standard_normal((n, 256)).astype(np.float32)
    vals, cols = oracle.topk(x, k, 2)
    sel = cols.astype(np.uint8)
    lhs = float((oracle.spgemm_fwd(indptr, indices, values, vals, sel).astype(np.float64) * g).sum())
    rhs = float((vals.astype(np.float64) * oracle.sspmm_bwd(indptr, indices, values, g, sel)).sum())
    assert abs(lhs - rhs) <= 1e-4 * max(1.0, abs(lhs))


def test_oracle_matches_reference_operator_glue():
    """ref_ops.npz: the reference's v3 / v4 / optimized autograd operators executed on the CPU
    (tests/golden/make_golden_ops.py).  Pins WHERE the degree divisions happen, which arrays the backward
    walks, and the v3 quirk of masking grad_output with the input's top-k pattern (spgemmfunction_v3.py:118)."""
    r = np.load(os.path.join(GOLD, "ref_ops.npz"))
    k = 32
    # v3: CSR forward / in_degrees, backward over the CSC arrays / out_degrees, grad_output masked first
    ip, ix, va, x, up = r["v3_indptr"], r["v3_indices"], r["v3_values"], r["v3_x"], r["v3_up"]
    n = len(ip) - 1
    import scipy.sparse as sp
    csc = sp.csr_matrix((va, ix, ip), shape=(n, n)).tocsc()
    csc.sort_indices()
    t_ptr, t_idx, t_val = csc.indptr.astype(np.int32), csc.indices.astype(np.int32), csc.data.astype(np.float32)
    in_deg = np.maximum(np.diff(ip), 1).astype(np.float32)
    out_deg = np.maximum(np.diff(t_ptr), 1).astype(np.float32)
    vals, cols = oracle.topk(x, k, 0)
    sel = cols.astype(np.uint8)
    np.testing.assert_allclose(oracle.spgemm_fwd(ip, ix, va, vals, sel, deg=in_deg), r["v3_out"], rtol=1e-5, atol=1e-6)
    mask = np.zeros_like(x)
    np.put_along_axis(mask, cols.astype(np.int64), 1.0, axis=1)
    gs = oracle.sspmm_bwd(t_ptr, t_idx, t_val, up * mask, sel, deg=out_deg)
    np.testing.assert_allclose(oracle.scatter_dense(gs, cols), r["v3_grad_input"], rtol=1e-5, atol=1e-6)
    # v4 / optimized: pre-computed torch.topk, undirected graph, one degree vector on both sides
    ip, ix, va, up = r["u_indptr"], r["u_indices"], r["u_values"], r["u_up"]
    tv, ti = r["u_topk_values"], r["u_topk_indices"]
    deg = np.maximum(np.diff(ip), 1).astype(np.float32)
    out = oracle.spgemm_fwd(ip, ix, va, tv, ti.astype(np.uint8), deg=deg)
    gs = oracle.sspmm_bwd(ip, ix, va, up, ti.astype(np.uint8), deg=deg)
    for name in ("v4", "opt"):
        np.testing.assert_allclose(out, r[name + "_out"], rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(gs, r[name + "_grad_topk_values"], rtol=1e-5, atol=1e-6)
