"""GPU: the reference-facing Python surface (autograd operators, MaxK nonlinearity, DirectMaxKKernels)
against the golden vectors produced by the reference's own code and against the oracle."""
import os

import numpy as np
import pytest
import torch

import oracle
from helpers import assert_close, graph_cuda, graph_np, make_problem

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def ref_py():
    return np.load(os.path.join(GOLD, "ref_py.npz"))


@pytest.fixture(scope="module")
def ref_cuda():
    return np.load(os.path.join(GOLD, "ref_cuda.npz"))


def _t(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a)).cuda()
    return t.to(dtype) if dtype is not None else t


# ---------------------------------------------------------------- golden: reference python operator
def test_v1_operator_matches_reference_forward(ref_py):
    """maxk_spgemm(...) == the reference's MaxKSpGEMMFunction.forward on the same inputs (its CPU branch)."""
    from maxk_spgemm_function import MaxKSpmmWrapper, maxk_spgemm
    ip, ix, va = _t(ref_py["indptr"]), _t(ref_py["indices"]), _t(ref_py["values"])
    x, deg, k = _t(ref_py["x"]), _t(ref_py["in_deg"]), int(ref_py["k"])
    w = MaxKSpmmWrapper("golden")
    w.build_metadata(ip)
    out = w.spmm(ix, va, x, k, in_degrees=deg)                       # warp4-driven, fused /in_degrees
    assert_close(out, ref_py["spgemm_fwd_norm"], "v1 forward / in_degrees")
    raw = maxk_spgemm(ix, va, x, k, None, 0, ip)                     # indptr-driven, no normalisation
    assert_close(raw, ref_py["spgemm_fwd_raw"], "v1 forward raw")
    assert w.num_warps * 4 == ref_py["warp4_small"].size
    assert np.array_equal(w.warp4_metadata.cpu().numpy(), ref_py["warp4_small"])   # == generate_meta.py output


@pytest.mark.parametrize("k", [8, 32])
def test_maxk_and_optmaxk_match_reference_classes(ref_py, k):
    from maxk_models_integrated import MaxK, OPTMaxK
    x = _t(ref_py["maxk_x"]).requires_grad_(True)
    up = _t(ref_py["maxk_up"])
    y = MaxK.apply(x, k)
    (g,) = torch.autograd.grad(y, x, up)
    assert np.array_equal(y.detach().cpu().numpy(), ref_py["maxk_fwd_k%d" % k])
    assert np.array_equal(g.cpu().numpy(), ref_py["maxk_bwd_k%d" % k])

    yo, tv, ti = OPTMaxK.apply(x, k)
    assert ti.dtype == torch.int64 and not ti.requires_grad
    assert np.array_equal(yo.detach().cpu().numpy(), ref_py["optmaxk_fwd_k%d" % k])
    # same (value, index) pairs as torch.topk, row order is ours
    order = np.argsort(-tv.detach().cpu().numpy(), axis=1, kind="stable")
    assert np.array_equal(np.take_along_axis(tv.detach().cpu().numpy(), order, 1), ref_py["optmaxk_vals_k%d" % k])
    assert np.array_equal(np.take_along_axis(ti.cpu().numpy(), order, 1), ref_py["optmaxk_idx_k%d" % k])
    up_v_sorted = ref_py["optmaxk_upv_k%d" % k]
    inv = np.argsort(order, axis=1)
    up_v = _t(np.take_along_axis(up_v_sorted, inv, 1))                # same upstream grads, permuted to our row order
    OPTMaxK.reference_compat = True
    try:
        (gc,) = torch.autograd.grad([yo, tv], x, [up, up_v], retain_graph=True)
        assert np.array_equal(gc.cpu().numpy(), ref_py["optmaxk_bwd_k%d" % k])     # reference drops grad_topk_values
    finally:
        OPTMaxK.reference_compat = False
    (gf,) = torch.autograd.grad([yo, tv], x, [up, up_v])
    full = ref_py["optmaxk_bwd_k%d" % k] + oracle.scatter_dense(up_v_sorted, ref_py["optmaxk_idx_k%d" % k])
    assert_close(gf, full, "OPTMaxK backward incl. grad_topk_values")


@pytest.mark.parametrize("k", [8, 32])
def test_optmaxk_value_desc_order_is_the_reference_output(ref_py, k):
    """OPTMaxK.order = "value_desc": (topk_values, topk_indices) are torch.topk's, element for element, as the
    reference class returns them (model_integrated_v3.py:32); uint8_indices=True hands out the kernel's selectors."""
    from maxk_models_integrated import OPTMaxK
    x = _t(ref_py["maxk_x"])
    OPTMaxK.order = "value_desc"
    try:
        yo, tv, ti = OPTMaxK.apply(x, k)
        _, tv8, ti8 = OPTMaxK.apply(x, k, True)
    finally:
        OPTMaxK.order = "banked"
    assert ti.dtype == torch.int64 and ti8.dtype == torch.uint8
    assert np.array_equal(tv.cpu().numpy(), ref_py["optmaxk_vals_k%d" % k])
    assert np.array_equal(ti.cpu().numpy(), ref_py["optmaxk_idx_k%d" % k])
    assert np.array_equal(ti8.cpu().numpy().astype(np.int64), ref_py["optmaxk_idx_k%d" % k]) and torch.equal(tv8, tv)
    assert np.array_equal(yo.cpu().numpy(), ref_py["optmaxk_fwd_k%d" % k])


# ------------------------------------------------------------------- golden: reference CUDA kernels
@pytest.mark.parametrize("name", ["k32", "k64"])
def test_entry_points_match_reference_cuda_kernels(ref_cuda, name):
    import maxk_cuda_kernels as kern
    g = {key[len(name) + 1:]: ref_cuda[key] for key in ref_cuda.files if key.startswith(name + "_")}
    k = g["data"].shape[1]
    w4 = _t(g["warp4"])
    out = kern.spmm_maxk_forward(w4, _t(g["indices"]), _t(g["values"]), _t(g["data"]), _t(g["sel"]), w4.numel() // 4, k)
    assert_close(out, g["fwd"], "spmm_maxk_forward vs reference kernel")
    gs = kern.spmm_maxk_backward(w4, _t(g["indices"]), _t(g["values"]), _t(g["grad"]), _t(g["sel"]), w4.numel() // 4, k)
    assert_close(gs, g["bwd"], "spmm_maxk_backward vs reference kernel", rtol=2e-5)


# ------------------------------------------------------------------------------ autograd end to end
@pytest.mark.parametrize("k", [16, 32])
def test_v1_autograd_forward_backward(k):
    from maxk_spgemm_function import MaxKSpmmWrapper
    p = make_problem(800, 30000, k, kind="powerlaw", seed=k, signed=True)
    ip, ix, va = graph_np(p["graph"])
    cip, cix, cva = graph_cuda(p["graph"])
    deg = np.maximum(np.diff(ip), 1).astype(np.float32)
    d = _t(deg)
    x = p["x"].cuda().requires_grad_(True)
    w = MaxKSpmmWrapper()
    w.build_metadata(cip)
    out = w.spmm(cix, cva, x, k, cip, d, d)
    up = p["grad"].cuda()
    out.backward(up)
    assert_close(out, oracle.spgemm_fwd(ip, ix, va, p["cbsr_val"], p["cbsr_sel"], deg=deg), "v1 fwd")
    gs = oracle.sspmm_bwd(ip, ix, va, p["grad"].numpy(), p["cbsr_sel"], deg=deg)
    assert_close(x.grad, oracle.scatter_dense(gs, p["cbsr_col"]), "v1 grad wrt input_features")
    assert x.grad.shape == (800, 256)


def test_v1_k_not_smaller_than_dim_keeps_every_feature():
    from maxk_spgemm_function import maxk_spgemm
    p = make_problem(300, 4000, 8, seed=1)
    ip, ix, va = graph_np(p["graph"])
    cip, cix, cva = graph_cuda(p["graph"])
    x = p["x"].cuda()
    out = maxk_spgemm(cix, cva, x, 256, None, 0, cip)
    ident = np.tile(np.arange(256, dtype=np.uint8), (300, 1))
    assert_close(out, oracle.spgemm_fwd(ip, ix, va, p["x"].numpy(), ident), "k == D", rtol=2e-5)


def test_v4_operator_with_precomputed_topk():
    from maxk_models_integrated import OPTMaxK
    from spgemmfunction_v4 import MaxKSpmmWrapper
    k = 32
    p = make_problem(600, 25000, k, kind="uniform", seed=4, signed=True)
    ip, ix, va = graph_np(p["graph"])
    cip, cix, cva = graph_cuda(p["graph"])
    deg = np.maximum(np.diff(ip), 1).astype(np.float32)
    x = p["x"].cuda().requires_grad_(True)
    _, tv, ti = OPTMaxK.apply(x, k)
    w = MaxKSpmmWrapper("g")
    w.build_metadata(cip)
    out = w.spmm(cix, cva, tv, ti, cip, _t(deg))
    assert_close(out, oracle.spgemm_fwd(ip, ix, va, p["cbsr_val"], p["cbsr_sel"], deg=deg), "v4 fwd")
    out.backward(p["grad"].cuda())
    gs = oracle.sspmm_bwd(ip, ix, va, p["grad"].numpy(), p["cbsr_sel"], deg=deg)
    assert_close(x.grad, oracle.scatter_dense(gs, p["cbsr_col"]), "v4 grad reaches x through OPTMaxK")
    # torch.topk output (int64, value order) is accepted as well
    tv2, ti2 = torch.topk(p["x"].cuda(), k, dim=1)
    out2 = w.spmm(cix, cva, tv2, ti2, cip, _t(deg))
    assert_close(out2, out.detach().cpu().numpy(), "v4 fwd with torch.topk inputs", rtol=2e-5)


def test_v3_operator_csr_forward_csc_backward(tmp_path, monkeypatch):
    """spgemmfunction_v3.py: CSR quads forward, CSC arrays + CSC quads backward, on a *directed* graph so that
    a mix-up of the two sets of arrays cannot go unnoticed; metadata read from the files generate_meta writes."""
    from graph_loader import GraphDataLoader, csr_to_csc, generate_meta
    from spgemmfunction_v3 import MaxKSpGEMMFunction, MaxKSpmmWrapper
    k = 16
    p = make_problem(500, 20000, k, kind="powerlaw", seed=9, signed=True)
    ip, ix, va = graph_np(p["graph"])
    cip, cix, cva = graph_cuda(p["graph"])
    t_ptr, t_idx, t_val = csr_to_csc(cip, cix, cva)
    in_deg = np.maximum(np.diff(ip), 1).astype(np.float32)
    out_deg = np.maximum(np.diff(t_ptr.cpu().numpy()), 1).astype(np.float32)
    monkeypatch.chdir(tmp_path)
    GraphDataLoader("kernels/graphs/").save_graph("toy", ip, ix)
    generate_meta("toy")
    w = MaxKSpmmWrapper("toy")
    with pytest.raises(RuntimeError):
        MaxKSpmmWrapper("absent").load_metadata()
    assert w.load_metadata() and w.num_warps_csr > 0 and w.num_warps_csc > 0
    x = p["x"].cuda().requires_grad_(True)
    out = w.spmm(cix, cva, x, k, cip, _t(in_deg), _t(out_deg), t_idx, t_val)
    assert_close(out, oracle.spgemm_fwd(ip, ix, va, p["cbsr_val"], p["cbsr_sel"], deg=in_deg), "v3 fwd")
    out.backward(p["grad"].cuda())
    # the kernel's formula on the CSC arrays (rows of A^T), gradient rows divided by out_degrees
    tp, ti_, tv_ = t_ptr.cpu().numpy(), t_idx.cpu().numpy(), t_val.cpu().numpy()
    gs = oracle.sspmm_bwd(tp, ti_, tv_, p["grad"].numpy(), p["cbsr_sel"], deg=out_deg)
    assert_close(x.grad, oracle.scatter_dense(gs, p["cbsr_col"]), "v3 grad (CSC arrays + CSC quads)")
    # reference_compat reproduces :118-119 (grad_output masked by the input's top-k pattern)
    MaxKSpGEMMFunction.reference_compat = True
    try:
        x2 = p["x"].cuda().requires_grad_(True)
        w.spmm(cix, cva, x2, k, cip, _t(in_deg), _t(out_deg), t_idx, t_val).backward(p["grad"].cuda())
    finally:
        MaxKSpGEMMFunction.reference_compat = False
    mask = np.zeros((500, 256), np.float32)
    np.put_along_axis(mask, p["cbsr_col"].astype(np.int64), 1.0, axis=1)
    gs2 = oracle.sspmm_bwd(tp, ti_, tv_, p["grad"].numpy() * mask, p["cbsr_sel"], deg=out_deg)
    assert_close(x2.grad, oracle.scatter_dense(gs2, p["cbsr_col"]), "v3 grad, reference_compat")


def test_optimized_operator_matches_v4_on_an_undirected_graph():
    """spgemmfunction.py: pre-computed top-k, CSC arrays with the CSR quads; on a symmetric graph (same row
    order either way) it is the v4 result."""
    from graph_loader import csr_to_csc
    from maxk_models_integrated import OPTMaxK
    from spgemmfunction import OptimizedMaxKSpmmWrapper
    from spgemmfunction_v4 import MaxKSpmmWrapper as V4
    from synth_graphs import symmetrize, synth_graph
    k = 32
    g = symmetrize(synth_graph(400, 6000, seed=11))
    cip, cix, cva = g["indptr"].cuda(), g["indices"].cuda(), g["values"].cuda()
    t_ptr, t_idx, t_val = csr_to_csc(cip, cix, cva)
    assert torch.equal(t_ptr, cip) and torch.equal(t_idx, cix)
    deg = (cip[1:] - cip[:-1]).clamp(min=1).float()
    x = torch.randn(400, 256, generator=torch.Generator().manual_seed(2)).cuda()
    up = torch.rand(400, 256, generator=torch.Generator().manual_seed(3)).cuda()
    res = []
    for make in (lambda: OptimizedMaxKSpmmWrapper("g"), lambda: V4("g")):
        xi = x.clone().requires_grad_(True)
        _, tv, ti = OPTMaxK.apply(xi, k)
        w = make()
        w.build_metadata(cip)
        out = (w.spmm(cix, cva, tv, ti, cip, deg, deg, t_idx, t_val) if isinstance(w, OptimizedMaxKSpmmWrapper)
               else w.spmm(cix, cva, tv, ti, cip, deg))
        out.backward(up)
        res.append((out.detach(), xi.grad))
    assert torch.equal(res[0][0], res[1][0])
    assert torch.allclose(res[0][1], res[1][1], rtol=1e-5, atol=1e-6)
    with pytest.raises(RuntimeError):
        OptimizedMaxKSpmmWrapper("g").spmm(cix, cva, x[:, :k], None, cip, deg, deg, t_idx, t_val)


def test_operators_match_reference_operator_glue():
    """ref_ops.npz = the reference's v3 / v4 / optimized autograd operators run on the CPU
    (tests/golden/make_golden_ops.py); ours on the GPU, same inputs, same call signatures."""
    import spgemmfunction as opt_mod
    import spgemmfunction_v3 as v3_mod
    import spgemmfunction_v4 as v4_mod
    from graph_loader import csr_to_csc
    import maxk_cuda_kernels as kern
    r = np.load(os.path.join(GOLD, "ref_ops.npz"))
    k = 32
    # v3 (13 arguments), with the reference's grad_output masking switched on
    ip, ix, va = _t(r["v3_indptr"]), _t(r["v3_indices"]), _t(r["v3_values"])
    t_ptr, t_idx, t_val = csr_to_csc(ip, ix, va)
    in_deg = (ip[1:] - ip[:-1]).clamp(min=1).float()
    out_deg = (t_ptr[1:] - t_ptr[:-1]).clamp(min=1).float()
    w_csr, n_csr = kern.build_warp4(ip)
    w_csc, n_csc = kern.build_warp4(t_ptr)
    x = _t(r["v3_x"]).requires_grad_(True)
    v3_mod.MaxKSpGEMMFunction.reference_compat = True
    try:
        y = v3_mod.maxk_spgemm(ix, va, x, k, w_csr, n_csr, ip, in_deg, out_deg, t_idx, t_val, w_csc, n_csc)
        y.backward(_t(r["v3_up"]))
    finally:
        v3_mod.MaxKSpGEMMFunction.reference_compat = False
    assert_close(y, r["v3_out"], "v3 forward vs the reference operator")
    assert_close(x.grad, r["v3_grad_input"], "v3 grad_input vs the reference operator")
    # v4 (8 arguments) and optimized (11 arguments): torch.topk output passed in as the reference's callers do
    ip, ix, va = _t(r["u_indptr"]), _t(r["u_indices"]), _t(r["u_values"])
    deg = (ip[1:] - ip[:-1]).clamp(min=1).float()
    w, nw = kern.build_warp4(ip)
    t_ptr, t_idx, t_val = csr_to_csc(ip, ix, va)
    ti = _t(r["u_topk_indices"])
    tv = _t(r["u_topk_values"]).requires_grad_(True)
    y4 = v4_mod.maxk_spgemm(ix, va, tv, ti, w, nw, ip, deg)
    y4.backward(_t(r["u_up"]))
    assert_close(y4, r["v4_out"], "v4 forward vs the reference operator")
    assert_close(tv.grad, r["v4_grad_topk_values"], "v4 grad_topk_values vs the reference operator")
    tv = _t(r["u_topk_values"]).requires_grad_(True)
    yo = opt_mod.optimized_maxk_spgemm(ix, va, tv, ti, w, nw, ip, deg, deg, t_idx, t_val)
    yo.backward(_t(r["u_up"]))
    assert_close(yo, r["opt_out"], "optimized forward vs the reference operator")
    assert_close(tv.grad, r["opt_grad_topk_values"], "optimized grad_topk_values vs the reference operator")


def test_fused_layer_entry_points():
    """maxk_layer_forward / maxk_layer_backward (additive) against the oracle, with the fused divisor."""
    import maxk_cuda_kernels as kern
    k = 32
    p = make_problem(900, 40000, k, kind="powerlaw", seed=12, signed=True)
    ip, ix, va = graph_np(p["graph"])
    cip, cix, cva = graph_cuda(p["graph"])
    deg = np.maximum(np.diff(ip), 1).astype(np.float32)
    out, vals, sel, masked = kern.maxk_layer_forward(cip, cix, cva, p["x"].cuda(), k, row_div=_t(deg), want_masked=True)
    assert np.array_equal(sel.cpu().numpy(), p["cbsr_sel"]) and np.array_equal(vals.cpu().numpy(), p["cbsr_val"])
    assert np.array_equal(masked.cpu().numpy(), oracle.maxk_act_fwd(p["x"].numpy(), p["cbsr_col"]))
    assert_close(out, oracle.spgemm_fwd(ip, ix, va, p["cbsr_val"], p["cbsr_sel"], deg=deg), "layer forward")
    gs, dense = kern.maxk_layer_backward(cip, cix, cva, p["grad"].cuda(), sel, row_div=_t(deg), dense_dim=256)
    want = oracle.sspmm_bwd(ip, ix, va, p["grad"].numpy(), p["cbsr_sel"], deg=deg)
    assert_close(gs, want, "layer backward")
    assert_close(dense, oracle.scatter_dense(want, p["cbsr_col"]), "layer backward, dense")
    assert kern.maxk_layer_backward(cip, cix, cva, p["grad"].cuda(), sel).shape == (900, k)


def test_gradcheck_like_adjoint_identity():
    """<fwd(x_vals), g> == <x_vals, bwd(g)> on the GPU kernels themselves."""
    import maxk_cuda_kernels as kern
    p = make_problem(700, 40000, 32, kind="powerlaw", seed=8, signed=True)
    cip, cix, cva = graph_cuda(p["graph"])
    data, sel, g = _t(p["cbsr_val"]), _t(p["cbsr_sel"]), p["grad"].cuda()
    out = kern.spgemm_forward_csr(cip[:-1], cip[1:], cix, cva, data, sel)
    gs = kern.sspmm_backward_csr(cip[:-1], cip[1:], cix, cva, g, sel)
    lhs, rhs = float((out.double() * g.double()).sum()), float((data.double() * gs.double()).sum())
    assert abs(lhs - rhs) <= 1e-5 * abs(lhs)


# ------------------------------------------------------------------------------ DirectMaxKKernels
def test_direct_kernel_interface(tmp_path, monkeypatch):
    from direct_kernel_interface import DirectMaxKKernels
    from graph_loader import GraphDataLoader, save_warp4, warp4_path
    p = make_problem(500, 20000, 32, seed=6)
    monkeypatch.chdir(tmp_path)
    loader = GraphDataLoader("kernels/graphs/")
    loader.save_graph("toy", p["graph"]["indptr"], p["graph"]["indices"])
    gd = loader.to_cuda_tensors(loader.load_graph("toy"))
    dk = DirectMaxKKernels("toy")
    assert dk.load_warp4_metadata() is False                          # no file yet, like the reference
    with pytest.raises(RuntimeError):
        dk.run_forward_kernel(gd, p["x"].cuda(), 32, timing=False)
    quads, _ = oracle.warp4(p["graph"]["indptr"].numpy(), 64)
    save_warp4(warp4_path("toy"), quads)
    assert dk.load_warp4_metadata() is True and dk.num_warps == quads.size // 4
    x = p["x"].cuda()
    for cuda_topk in (True, False):
        assert dk.validate_against_cusparse(gd, x, 32, use_cuda_topk=cuda_topk)
        assert dk.last_validation["max_error"] < 1e-4
    out, ms = dk.run_forward_kernel(gd, x, 32, timing=True)
    vals, cols = oracle.topk(p["x"].numpy(), 32, 0)
    assert_close(out, oracle.spgemm_fwd(gd["indptr"].cpu().numpy(), gd["indices"].cpu().numpy(), gd["values"].cpu().numpy(),
                                        vals, cols.astype(np.uint8)), "DirectMaxKKernels forward")
    assert ms > 0
    gi, ms_b = dk.run_backward_kernel(gd, p["grad"].cuda(), 32, timing=True)
    assert gi.shape == (500, 32) and ms_b > 0
    res = dk.benchmark_all_k_values(gd, k_values=[16, 32, 64, 96], verbose=False)
    assert sorted(res) == [16, 32, 64] and all(v["forward_time"] > 0 for v in res.values())
    dk2 = DirectMaxKKernels("toy")
    dk2.build_warp4_metadata(gd)
    assert torch.equal(dk2.warp4_metadata, dk.warp4_metadata)


def test_generate_meta_writes_csr_and_csc_quads(tmp_path, monkeypatch):
    """kernels/generate_meta.py + generate_meta_csc.py on the GPU: both files equal the oracle's quads."""
    from graph_loader import GraphDataLoader, csr_to_csc, generate_meta
    p = make_problem(700, 30000, 32, seed=8, kind="powerlaw")
    monkeypatch.chdir(tmp_path)
    GraphDataLoader("kernels/graphs/").save_graph("toy", p["graph"]["indptr"], p["graph"]["indices"])
    p_csr, n_csr, p_csc, n_csc = generate_meta("toy")
    ip = p["graph"]["indptr"].numpy()
    want, w = oracle.warp4(ip, 64)
    assert n_csr == w and np.array_equal(np.fromfile(p_csr, dtype=np.int32), want)
    t_ptr, _ = csr_to_csc(p["graph"]["indptr"], p["graph"]["indices"])
    want_t, w_t = oracle.warp4(t_ptr.numpy(), 64)
    assert n_csc == w_t and np.array_equal(np.fromfile(p_csc, dtype=np.int32), want_t)
    assert p_csc.endswith("w12_nz64_warp_4_csc/toy.warp4_csc")


def test_reference_smoke_shape_and_bug_repro_ks():
    """V=1000, E=5000 random graph (maxk_spgemm_function.py:279-286) and the k values of test_bug.py:16-20:
    the reference's uint8 top-k faults for k in {8,16,18}; ours is exact for every k."""
    import maxk_cuda_kernels as kern
    torch.manual_seed(42)
    x = torch.rand(3349, 256, device="cuda")
    for k in (8, 16, 18, 19, 20, 32):
        v, i = kern.cuda_topk_maxk_float(x, k)
        tv, ti = torch.topk(x, k, dim=1)
        assert torch.equal(v, tv) and torch.equal(x.gather(1, i.long()), v)
        tie_free = (tv[:, 1:] != tv[:, :-1]).all(dim=1)       # torch.rand has 2^-24 granularity: a few rows tie
        assert tie_free.float().mean() > 0.9 and torch.equal(i.long()[tie_free], ti[tie_free])
    xu = (x * 255).round().to(torch.uint8)
    v8, i8 = kern.cuda_topk_maxk(xu, 32)
    ev, ec = oracle.topk(xu.float().cpu().numpy(), 32, 0)
    assert v8.dtype == torch.uint8 and np.array_equal(v8.cpu().numpy(), ev.astype(np.uint8))
    assert np.array_equal(i8.cpu().numpy(), ec.astype(np.uint8))
    sel = kern.generate_sparse_selector(100, 256, 32)
    assert sel.shape == (100, 32) and sel.dtype == torch.uint8
    assert all(len(set(r.tolist())) == 32 for r in sel.cpu())


def test_host_staged_pipeline_matches_device_resident_path():
    """maxk_host_pipeline: slab-wise overlapped copies + kernels give the same results as one-shot calls."""
    import maxk_cuda_kernels as kern
    from maxk_host_pipeline import HostStagedMaxKLayer
    p = make_problem(3000, 90000, 32, kind="powerlaw", seed=12, signed=True)
    ip, ix, va = graph_np(p["graph"])
    cip, cix, cva = graph_cuda(p["graph"])
    hx, hg = p["x"].pin_memory(), p["grad"].pin_memory()
    hout, hgs = torch.empty(3000, 256).pin_memory(), torch.empty(3000, 32).pin_memory()
    layer = HostStagedMaxKLayer(cip, cix, cva, 32, slabs=5)
    for _ in range(3):                                   # repeated calls reuse the double-buffered staging
        layer.run(hx, hg, hout, hgs, block_current_stream=False)
    torch.cuda.synchronize()
    assert_close(hout, oracle.spgemm_fwd(ip, ix, va, p["cbsr_val"], p["cbsr_sel"]), "staged forward")
    assert_close(hgs, oracle.sspmm_bwd(ip, ix, va, p["grad"].numpy(), p["cbsr_sel"]), "staged backward", rtol=2e-5)
    ref = kern.spgemm_forward_csr(cip[:-1], cip[1:], cix, cva, _t(p["cbsr_val"]), _t(p["cbsr_sel"]))
    assert torch.equal(hout.cuda(), ref)                 # the forward is deterministic, slab by slab too


def test_layer_is_cuda_graph_capturable():
    """top-k + forward + backward captured once in a CUDA graph and replayed on new inputs (no syncs, no
    host copies, no default-stream work inside the calls)."""
    import maxk_cuda_kernels as kern
    p = make_problem(2000, 60000, 32, kind="powerlaw", seed=14, signed=True)
    ip, ix, va = graph_np(p["graph"])
    cip, cix, cva = graph_cuda(p["graph"])
    x, g = p["x"].cuda(), p["grad"].cuda()
    out, gs = torch.empty(2000, 256, device="cuda"), torch.empty(2000, 32, device="cuda")
    vals, sel = torch.empty(2000, 32, device="cuda"), torch.empty(2000, 32, device="cuda", dtype=torch.uint8)

    def step():
        kern.topk_cbsr(x, 32, out_values=vals, out_sel=sel)
        kern.spgemm_forward_csr(cip[:-1], cip[1:], cix, cva, vals, sel, out=out)
        kern.sspmm_backward_csr(cip[:-1], cip[1:], cix, cva, g, sel, out=gs)

    step()                                            # warm-up outside the capture (function attributes)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        step()
    x.copy_(torch.randn(2000, 256, generator=torch.Generator().manual_seed(99)))   # new inputs, same buffers
    g.copy_(torch.rand(2000, 256, generator=torch.Generator().manual_seed(98)))
    out.zero_(); gs.zero_()
    graph.replay()
    torch.cuda.synchronize()
    v2, c2 = oracle.topk(x.cpu().numpy(), 32, 2)
    assert np.array_equal(sel.cpu().numpy(), c2.astype(np.uint8))
    assert_close(out, oracle.spgemm_fwd(ip, ix, va, v2, c2.astype(np.uint8)), "graph replay forward")
    assert_close(gs, oracle.sspmm_bwd(ip, ix, va, g.cpu().numpy(), c2.astype(np.uint8)), "graph replay backward", rtol=2e-5)


def test_ops_run_on_the_current_stream_and_device():
    """Launches go to torch's CURRENT stream (the reference uses the legacy default stream,
    cuda_kernel_wrappers.cu:46) and to the tensors' device."""
    import maxk_cuda_kernels as kern
    p = make_problem(1500, 50000, 32, seed=21)
    ip, ix, va = graph_np(p["graph"])
    cip, cix, cva = graph_cuda(p["graph"])
    side = torch.cuda.Stream()
    x = p["x"].cuda()
    torch.cuda.synchronize()
    with torch.cuda.stream(side):
        big = torch.empty(64 << 20, device="cuda").normal_()      # keeps `side` busy before our kernels
        r = kern.topk_cbsr(x, 32)
        out = kern.spgemm_forward_csr(cip[:-1], cip[1:], cix, cva, r["values"], r["sel"])
        gs = kern.sspmm_backward_csr(cip[:-1], cip[1:], cix, cva, p["grad"].cuda(non_blocking=True), r["sel"])
    side.synchronize()
    assert_close(out, oracle.spgemm_fwd(ip, ix, va, p["cbsr_val"], p["cbsr_sel"]), "forward on a side stream")
    assert_close(gs, oracle.sspmm_bwd(ip, ix, va, p["grad"].numpy(), p["cbsr_sel"]), "backward on a side stream")
    del big
    if torch.cuda.device_count() > 1:                              # tensors on cuda:1 while cuda:0 is current
        d1 = torch.device("cuda", 1)
        r1 = kern.topk_cbsr(x.to(d1), 32)
        o1 = kern.spgemm_forward_csr(cip[:-1].to(d1), cip[1:].to(d1), cix.to(d1), cva.to(d1), r1["values"], r1["sel"])
        torch.cuda.synchronize(d1)
        assert o1.device == d1 and torch.equal(o1.cpu(), out.cpu())
