"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle on the same seeded inputs.

Bar (BASELINE.json north star): top-k index sets bit-exact with lowest-column tie-breaking;
fp32 outputs within rtol 1e-5 / atol 1e-6 of the fp64 oracle.
"""
import numpy as np
import pytest
import torch

import oracle
from helpers import assert_close, graph_cuda, graph_np, make_problem

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def kern():
    import maxk_cuda_kernels
    return maxk_cuda_kernels


# ------------------------------------------------------------------------------------------ top-k
@pytest.mark.parametrize("k", [1, 8, 16, 19, 32, 64, 100, 256])
@pytest.mark.parametrize("signed", [False, True])
def test_topk_exact(kern, k, signed):
    gen = torch.Generator().manual_seed(k)
    x = torch.randn(517, 256, generator=gen) if signed else torch.rand(517, 256, generator=gen)
    for order in (0, 1, 2):
        ev, ec = oracle.topk(x.numpy(), k, order)
        r = kern.topk_cbsr(x.cuda(), k, order=order, want_i32=True, want_i64=True, want_masked=True)
        assert np.array_equal(r["sel"].cpu().numpy(), ec.astype(np.uint8))
        assert np.array_equal(r["i32"].cpu().numpy(), ec)
        assert np.array_equal(r["i64"].cpu().numpy(), ec.astype(np.int64))
        assert np.array_equal(r["values"].cpu().numpy(), ev)       # selected values are copied bit-exactly
        assert np.array_equal(r["masked"].cpu().numpy(), oracle.maxk_act_fwd(x.numpy(), ec))


def test_topk_ties_lowest_column(kern):
    """Adversarial ties straddling rank k: the lowest columns must win (torch.topk does not promise this)."""
    rows = [
        [1, 3, 3, 3, 2, 3, 0, 3] * 32,                 # SURVEY.md hard part 3 probe, tiled to 256
        [5.0] * 256,                                    # all equal
        [0.0, -0.0] * 128,                              # -0 == +0
        [float(i % 7) for i in range(256)],
        [float("nan") if i in (200, 3) else float(i % 5) for i in range(256)],   # NaN is the largest
        [float("inf") if i % 50 == 0 else -float("inf") for i in range(256)],
    ]
    x = torch.tensor(rows, dtype=torch.float32)
    for k in (1, 2, 5, 8, 16, 32, 33, 64, 128, 255):
        for order in (0, 1, 2):
            ev, ec = oracle.topk(x.numpy(), k, order)
            r = kern.topk_cbsr(x.cuda(), k, order=order)
            assert np.array_equal(r["sel"].cpu().numpy(), ec.astype(np.uint8)), (k, order)
            got = r["values"].cpu().numpy()
            assert np.array_equal(np.isnan(got), np.isnan(ev)) and np.array_equal(np.nan_to_num(got), np.nan_to_num(ev))


@pytest.mark.parametrize("k", [8, 16, 32, 64])
def test_topk_general_and_specialised_kernels_agree(kern, k):
    """dim 256 / banked order / k in {8,16,32,64} runs the specialised kernel when the rows are 32-byte aligned
    and the general one otherwise (include/maxk_b200.h): a feature matrix that starts 16 bytes into a buffer
    must give bit-identical CBSR, indices and masked rows."""
    gen = torch.Generator().manual_seed(100 + k)
    x = torch.randn(1031, 256, generator=gen)
    x[5] = 1.0                                       # ties straddling rank k
    x[6, ::3] = float("nan")
    x[7] = torch.tensor([0.0, -0.0] * 128)
    buf = torch.empty(1031 * 256 + 4, device="cuda")
    shifted = buf[4:].view(1031, 256)
    shifted.copy_(x)
    aligned = x.cuda()
    assert aligned.data_ptr() % 32 == 0 and shifted.data_ptr() % 32 == 16
    a = kern.topk_cbsr(aligned, k, order=2, want_i32=True, want_masked=True)
    b = kern.topk_cbsr(shifted, k, order=2, want_i32=True, want_masked=True)
    ev, ec = oracle.topk(x.numpy(), k, 2)
    for r in (a, b):
        assert np.array_equal(r["sel"].cpu().numpy(), ec.astype(np.uint8))
        assert np.array_equal(r["i32"].cpu().numpy(), ec)
    assert torch.equal(a["values"].view(torch.int32), b["values"].view(torch.int32))
    assert torch.equal(a["masked"].view(torch.int32), b["masked"].view(torch.int32))


@pytest.mark.parametrize("dim,k", [(64, 8), (100, 19), (255, 32), (7, 7), (1, 1)])
def test_topk_other_dims(kern, dim, k):
    x = torch.randn(333, dim, generator=torch.Generator().manual_seed(dim))
    ev, ec = oracle.topk(x.numpy(), k, 0)
    r = kern.topk_cbsr(x.cuda(), k, order=0, want_masked=True)
    assert np.array_equal(r["sel"].cpu().numpy(), ec.astype(np.uint8))
    assert np.array_equal(r["values"].cpu().numpy(), ev)
    assert np.array_equal(r["masked"].cpu().numpy(), oracle.maxk_act_fwd(x.numpy(), ec))


def test_topk_matches_torch_topk_on_tie_free_rows(kern):
    x = torch.rand(2000, 256, generator=torch.Generator().manual_seed(7)).cuda()
    v, i = kern.cuda_topk_maxk_float(x, 32)
    tv, ti = torch.topk(x, 32, dim=1)
    tie_free = (tv[:, 1:] != tv[:, :-1]).all(dim=1)
    assert torch.equal(v, tv) and torch.equal(i.long()[tie_free], ti[tie_free]) and tie_free.float().mean() > 0.9
    assert i.dtype == torch.int32


# ------------------------------------------------------------------------------------ fwd / bwd
CASES = [
    # n, e, kind
    (300, 3000, "uniform"),
    (1000, 5000, "uniform"),        # the reference's smoke shape (maxk_spgemm_function.py:279)
    (997, 60000, "powerlaw"),
    (64, 40000, "uniform"),         # dense-ish: long rows relative to n
]


@pytest.mark.parametrize("n,e,kind", CASES)
@pytest.mark.parametrize("k", [8, 16, 32, 64, 19, 3])
@pytest.mark.parametrize("order", [2, 0])     # kernels must be correct for ANY entry order, fastest on 2
def test_forward_backward_vs_oracle(kern, n, e, kind, k, order):
    p = make_problem(n, e, k, kind=kind, seed=k, signed=True, order=order)
    ip, ix, va = graph_np(p["graph"])
    cip, cix, cva = graph_cuda(p["graph"])
    data, sel = torch.from_numpy(p["cbsr_val"]).cuda(), torch.from_numpy(p["cbsr_sel"]).cuda()
    out = kern.spgemm_forward_csr(cip[:-1], cip[1:], cix, cva, data, sel)
    assert_close(out, oracle.spgemm_fwd(ip, ix, va, p["cbsr_val"], p["cbsr_sel"]), "forward")
    gs = kern.sspmm_backward_csr(cip[:-1], cip[1:], cix, cva, p["grad"].cuda(), sel)
    assert_close(gs, oracle.sspmm_bwd(ip, ix, va, p["grad"].numpy(), p["cbsr_sel"]), "backward")


def test_fused_degree_division(kern):
    p = make_problem(500, 20000, 32, kind="powerlaw", seed=3)
    ip, ix, va = graph_np(p["graph"])
    cip, cix, cva = graph_cuda(p["graph"])
    deg = np.maximum(np.diff(ip), 1).astype(np.float32)
    data, sel = torch.from_numpy(p["cbsr_val"]).cuda(), torch.from_numpy(p["cbsr_sel"]).cuda()
    d = torch.from_numpy(deg).cuda()
    out = kern.spgemm_forward_csr(cip[:-1], cip[1:], cix, cva, data, sel, row_div=d)
    assert_close(out, oracle.spgemm_fwd(ip, ix, va, p["cbsr_val"], p["cbsr_sel"], deg=deg), "forward/deg")
    gs = kern.sspmm_backward_csr(cip[:-1], cip[1:], cix, cva, p["grad"].cuda(), sel, row_div=d)
    assert_close(gs, oracle.sspmm_bwd(ip, ix, va, p["grad"].numpy(), p["cbsr_sel"], deg=deg), "backward/deg")


def test_long_rows_take_the_cta_path(kern):
    """Rows longer than kLongRow (4096 edges) are reduced by a whole CTA; empty rows stay zero."""
    n, k = 3000, 32
    deg = np.zeros(n, np.int64)
    deg[5], deg[77], deg[2999] = 5000, 20000, 4097
    deg[100:200] = 37
    indptr = np.zeros(n + 1, np.int32)
    indptr[1:] = np.cumsum(deg)
    rng = np.random.default_rng(1)
    indices = rng.integers(0, n, indptr[-1]).astype(np.int32)
    values = rng.random(indptr[-1], dtype=np.float32)
    x = rng.standard_normal((n, 256)).astype(np.float32)
    grad = rng.random((n, 256), dtype=np.float32)
    vals, cols = oracle.topk(x, k, 1)
    sel = cols.astype(np.uint8)
    cip, cix, cva = (torch.from_numpy(a).cuda() for a in (indptr, indices, values))
    out = kern.spgemm_forward_csr(cip[:-1], cip[1:], cix, cva, torch.from_numpy(vals).cuda(), torch.from_numpy(sel).cuda())
    exp = oracle.spgemm_fwd(indptr, indices, values, vals, sel)
    assert_close(out, exp, "forward long rows", rtol=2e-5)
    assert float(out[0].abs().max()) == 0.0 and float(out[2998].abs().max()) == 0.0
    gs = kern.sspmm_backward_csr(cip[:-1], cip[1:], cix, cva, torch.from_numpy(grad).cuda(), torch.from_numpy(sel).cuda())
    assert_close(gs, oracle.sspmm_bwd(indptr, indices, values, grad, sel), "backward long rows", rtol=2e-5)


def test_empty_graph_and_empty_rows(kern):
    n, k = 50, 32
    p = make_problem(n, 200, k, seed=9)
    sel, data = torch.from_numpy(p["cbsr_sel"]).cuda(), torch.from_numpy(p["cbsr_val"]).cuda()
    zero_ptr = torch.zeros(n + 1, dtype=torch.int32, device="cuda")
    e0i, e0v = torch.empty(0, dtype=torch.int32, device="cuda"), torch.empty(0, dtype=torch.float32, device="cuda")
    out = kern.spgemm_forward_csr(zero_ptr[:-1], zero_ptr[1:], e0i, e0v, data, sel)
    assert out.shape == (n, 256) and float(out.abs().max()) == 0.0
    gs = kern.sspmm_backward_csr(zero_ptr[:-1], zero_ptr[1:], e0i, e0v, p["grad"].cuda(), sel)
    assert gs.shape == (n, k) and float(gs.abs().max()) == 0.0


def test_forward_is_deterministic(kern):
    p = make_problem(2000, 100000, 32, kind="powerlaw", seed=5)
    cip, cix, cva = graph_cuda(p["graph"])
    data, sel = torch.from_numpy(p["cbsr_val"]).cuda(), torch.from_numpy(p["cbsr_sel"]).cuda()
    a = kern.spgemm_forward_csr(cip[:-1], cip[1:], cix, cva, data, sel)
    b = kern.spgemm_forward_csr(cip[:-1], cip[1:], cix, cva, data, sel)
    assert torch.equal(a, b)


# --------------------------------------------------------------------------------------- warp4
@pytest.mark.parametrize("n,e,kind", [(300, 3000, "uniform"), (5000, 400000, "powerlaw"), (10, 0, "uniform")])
def test_build_warp4_matches_oracle(kern, n, e, kind):
    if e == 0:
        indptr = np.zeros(n + 1, np.int32)
    else:
        indptr = make_problem(n, e, 8, kind=kind)["graph"]["indptr"].numpy()
    exp, w = oracle.warp4(indptr, 64)
    got, gw = kern.build_warp4(torch.from_numpy(indptr).cuda(), 64)
    assert gw == w and np.array_equal(got.cpu().numpy(), exp)


def test_reference_entry_points_driven_by_warp4(kern):
    """spmm_maxk_forward / spmm_maxk_backward receive warp4 quads and no indptr (cuda_kernel_bindings.cpp:42-50)."""
    p = make_problem(1200, 90000, 32, kind="powerlaw", seed=11)
    ip, ix, va = graph_np(p["graph"])
    cip, cix, cva = graph_cuda(p["graph"])
    w4, nw = kern.build_warp4(cip, 64)
    data, sel = torch.from_numpy(p["cbsr_val"]).cuda(), torch.from_numpy(p["cbsr_sel"]).cuda()
    out = kern.spmm_maxk_forward(w4, cix, cva, data, sel, nw, 32)
    assert out.shape == (1200, 256)
    assert_close(out, oracle.spgemm_fwd_warp4(w4.cpu().numpy(), ix, va, p["cbsr_val"], p["cbsr_sel"], 1200), "fwd/warp4")
    gs = kern.spmm_maxk_backward(w4, cix, cva, p["grad"].cuda(), sel, nw, 32)
    assert_close(gs, oracle.sspmm_bwd(ip, ix, va, p["grad"].numpy(), p["cbsr_sel"]), "bwd/warp4")


# ------------------------------------------------------------------------------- helper kernels
def test_scatter_mask_dense_spmm(kern):
    p = make_problem(400, 9000, 32, seed=2)
    ip, ix, va = graph_np(p["graph"])
    cip, cix, cva = graph_cuda(p["graph"])
    vals, sel = torch.from_numpy(p["cbsr_val"]).cuda(), torch.from_numpy(p["cbsr_sel"]).cuda()
    dense = kern.cbsr_scatter(vals, sel)
    assert np.array_equal(dense.cpu().numpy(), oracle.scatter_dense(p["cbsr_val"], p["cbsr_col"]))
    m = kern.mask_apply(p["grad"].cuda(), sel)
    assert np.array_equal(m.cpu().numpy(), oracle.maxk_act_bwd(p["grad"].numpy(), p["cbsr_col"]))
    m2 = kern.mask_apply(p["grad"].cuda(), sel, vals)
    assert_close(m2, oracle.maxk_act_bwd(p["grad"].numpy(), p["cbsr_col"]) + oracle.scatter_dense(p["cbsr_val"], p["cbsr_col"]), "mask+add")
    ref = kern.cusparse_spmm(cip, cix, cva, dense)
    assert_close(ref, oracle.spgemm_fwd(ip, ix, va, p["cbsr_val"], p["cbsr_sel"]), "dense spmm", rtol=2e-5)


def test_errors_are_loud(kern):
    x = torch.rand(8, 256)
    with pytest.raises(RuntimeError):
        kern.topk_cbsr(x, 32)                       # CPU tensor: no fallback
    with pytest.raises(RuntimeError):
        kern.topk_cbsr(x.cuda(), 0)
    with pytest.raises(RuntimeError):
        kern.topk_cbsr(torch.rand(8, 300).cuda(), 32)   # uint8 selectors cannot address 300 columns
    with pytest.raises(RuntimeError):
        kern.load_warp4_metadata("no_such_graph")


def test_c_abi_status_codes(kern):
    """Argument errors come back as negative status codes from the C ABI itself (no launch)."""
    import ctypes
    lib = kern._lib
    x = torch.rand(4, 256, device="cuda")
    vals = torch.empty(4, 8, device="cuda")
    P = lambda t: ctypes.c_void_p(t.data_ptr())
    null = ctypes.c_void_p(0)
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    assert lib.maxk_topk_cbsr(P(x), 4, 256, 0, 2, P(vals), null, null, null, null, st) == -1      # MAXK_ERR_BAD_K
    assert lib.maxk_topk_cbsr(P(x), 4, 300, 8, 2, P(vals), null, null, null, null, st) == -2      # MAXK_ERR_BAD_DIM
    assert lib.maxk_topk_cbsr(null, 4, 256, 8, 2, P(vals), null, null, null, null, st) == -3      # MAXK_ERR_NULL
    assert lib.maxk_topk_cbsr(P(x), 4, 256, 8, 7, P(vals), null, null, null, null, st) == -6      # bad order
    assert lib.maxk_topk_cbsr(P(x), 0, 256, 8, 2, P(vals), null, null, null, null, st) == 0       # empty input is fine
    ip = torch.zeros(5, dtype=torch.int32, device="cuda")
    sel = torch.zeros(4, 8, dtype=torch.uint8, device="cuda")
    out = torch.empty(4, 256, device="cuda")
    ws = torch.empty(4096, dtype=torch.uint8, device="cuda")
    args = (P(ip), ctypes.c_void_p(ip.data_ptr() + 4), null, null, P(vals), P(sel), P(out), 4, 0, 256, 8, null)
    assert lib.maxk_spgemm_forward(*args, P(ws), 8, st) == -4                                      # workspace too small
    assert lib.maxk_spgemm_forward(*args, ctypes.c_void_p(ws.data_ptr() + 4), 4000, st) == -5      # misaligned
    assert lib.maxk_spgemm_forward(*args, P(ws), 4096, st) == 0
    torch.cuda.synchronize()
    assert float(out.abs().max()) == 0.0
    assert "workspace" in lib.maxk_status_string(-4).decode()
