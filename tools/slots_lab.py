"""Developer tool (GPU): the slot-parallel forward (csrc/spgemm_fwd_slots.cu + plan.cu) against the round-1
row-per-warp forward on the BASELINE shapes: parity of the two and CUDA-event timings.

    python tools/slots_lab.py [--shapes reddit,yelp,...] [--quick]
"""
import argparse
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "spgemm-prunning_b200"), os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)

import torch  # noqa: E402
import maxk_cuda_kernels as K  # noqa: E402
from synth_graphs import SHAPES, synth_graph  # noqa: E402

lib = K._lib
c_i64, c_int, c_ptr, c_size = ctypes.c_int64, ctypes.c_int, ctypes.c_void_p, ctypes.c_size_t
lib.maxk_plan_bytes.restype = c_size
lib.maxk_plan_bytes.argtypes = [c_i64]
lib.maxk_plan_workspace_bytes.restype = c_size
lib.maxk_plan_workspace_bytes.argtypes = [c_i64]
lib.maxk_plan_build.restype = c_int
lib.maxk_plan_build.argtypes = [c_ptr, c_ptr, c_i64, c_ptr, c_size, c_ptr, c_size, c_ptr]
lib.maxk_spgemm_forward_planned.restype = c_int
lib.maxk_spgemm_forward_planned.argtypes = [c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_i64, c_i64, c_int, c_int, c_ptr, c_ptr]
lib.maxk_sspmm_backward_planned.restype = c_int
lib.maxk_sspmm_backward_planned.argtypes = [c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_i64, c_i64, c_i64, c_int, c_int, c_ptr, c_int, c_ptr]


def ptr(t):
    return c_ptr(t.data_ptr()) if t is not None else c_ptr(0)


def stream():
    return c_ptr(torch.cuda.current_stream().cuda_stream)


def build_plan(rb, re_):
    n = rb.numel()
    pb, wb = lib.maxk_plan_bytes(n), lib.maxk_plan_workspace_bytes(n)
    plan = torch.empty((pb + 15) // 16 * 16, dtype=torch.uint8, device=rb.device)
    ws = torch.empty((wb + 15) // 16 * 16, dtype=torch.uint8, device=rb.device)
    st = lib.maxk_plan_build(ptr(rb), ptr(re_), n, ptr(plan), pb, ptr(ws), wb, stream())
    assert st == 0, st
    return plan


def fwd_planned(plan, ix, va, vals, sel, n_rows, dim=256, row_div=None, out=None):
    if out is None:
        out = torch.empty(n_rows, dim, device=vals.device)
    st = lib.maxk_spgemm_forward_planned(ptr(plan), ptr(ix), ptr(va), ptr(vals), ptr(sel), ptr(out), n_rows, ix.numel(),
                                         dim, vals.size(1), ptr(row_div), stream())
    assert st == 0, st
    return out


def bwd_planned(plan, ix, va, grad, sel, n_rows, row_div=None, out=None):
    n_dst, k = sel.shape
    if out is None:
        out = torch.empty(n_dst, k, device=grad.device)
    st = lib.maxk_sspmm_backward_planned(ptr(plan), ptr(ix), ptr(va), ptr(grad), ptr(sel), ptr(out), n_rows, n_dst, ix.numel(),
                                         grad.size(1), k, ptr(row_div), 0, stream())
    assert st == 0, st
    return out


def slot_order(vals, sel):
    """CBSR rows (any order) -> the slot kernels' order: residue classes mod 4, largest class first (ties: lower
    class), columns ascending inside a class; sorted rank p -> lane p // EPL, instruction p % EPL."""
    n, k = sel.shape
    c = sel.long()
    cls = c & 3
    sizes = torch.stack([(cls == j).sum(1) for j in range(4)], 1)                     # [n, 4]
    order = torch.argsort(-sizes * 4 + torch.arange(4, device=c.device), dim=1, stable=True)   # classes by (size desc, class)
    rank = torch.empty_like(order)
    rank.scatter_(1, order, torch.arange(4, device=c.device).expand(n, 4))
    key = rank.gather(1, cls) * 256 + c
    perm = torch.argsort(key, dim=1)
    sv, sc = vals.gather(1, perm), sel.gather(1, perm)
    if k >= 32 and k % 32 == 0:
        epl = k // 4
        p = torch.arange(k, device=c.device)
        t, i = p // epl, p % epl
        mem = 32 * (i // 8) + 8 * t + (i % 8)
        ov, oc = torch.empty_like(sv), torch.empty_like(sc)
        ov[:, mem] = sv
        oc[:, mem] = sc
        return ov.contiguous(), oc.contiguous()
    return sv.contiguous(), sc.contiguous()


def time_ms(fn, warm=3, reps=8):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    for a, b in evs:
        a.record()
        fn()
        b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in evs)
    return ts[len(ts) // 2], ts[0]


def small_checks():
    import numpy as np
    import oracle
    torch.manual_seed(0)
    dev = "cuda"
    cases = []
    for (n, e, k, dim, kind) in [(300, 3000, 32, 256, "uniform"), (1000, 5000, 32, 256, "uniform"), (257, 40000, 32, 256, "powerlaw"),
                                 (64, 64 * 5000, 32, 256, "uniform"), (500, 100000, 16, 256, "powerlaw"), (500, 9000, 8, 256, "uniform"),
                                 (400, 70000, 64, 256, "uniform"), (300, 5000, 19, 256, "uniform"), (300, 5000, 96, 256, "uniform"),
                                 (300, 5000, 128, 256, "uniform"), (300, 6000, 32, 200, "uniform"), (300, 6000, 12, 100, "uniform"),
                                 (3, 0, 32, 256, "uniform")]:
        cases.append((n, e, k, dim, kind))
    for (n, e, k, dim, kind) in cases:
        g = synth_graph(n, e, seed=7, kind=kind) if e else {"indptr": torch.zeros(n + 1, dtype=torch.int32), "indices": torch.zeros(0, dtype=torch.int32), "values": torch.zeros(0)}
        ip, ix, va = g["indptr"].to(dev), g["indices"].to(dev), g["values"].to(dev)
        if n == 300 and k == 32 and dim == 256:      # a few empty rows
            pass
        x = torch.randn(n, dim, device=dev)
        r = K.topk_cbsr(x, k, order=K.ORDER_COLUMN_ASC)
        vals, sel = slot_order(r["values"], r["sel"])
        deg = torch.clamp((ip[1:] - ip[:-1]).float(), min=1.0)
        plan = build_plan(ip[:-1].contiguous(), ip[1:].contiguous())
        hdr = plan[:64].view(torch.int32).tolist()
        for div in (None, deg):
            out = fwd_planned(plan, ix, va, vals, sel, n, dim=dim, row_div=div)
            exp = oracle.spgemm_fwd(ip.cpu().numpy(), ix.cpu().numpy(), va.cpu().numpy(), vals.cpu().numpy(),
                                    sel.cpu().numpy(), dim=dim, deg=None if div is None else deg.cpu().numpy())
            got = out.cpu().numpy()
            err = np.abs(got - exp).max()
            tol = 1e-5 * np.abs(exp).max() + 1e-6
            ok = np.allclose(got, exp, rtol=2e-5, atol=1e-5)
            print("small n=%d e=%d k=%d dim=%d %s div=%s: max err %.3g %s  plan nA=%d nB=%d nC=%d nD=%d" % (
                n, e, k, dim, kind, div is not None, err, "OK" if ok else "MISMATCH", hdr[2], hdr[3], hdr[4], hdr[5]), flush=True)
            assert ok
            gr = torch.randn(n, dim, device=dev)
            gs = bwd_planned(plan, ix, va, gr, sel, n, row_div=div)
            exp = oracle.sspmm_bwd(ip.cpu().numpy(), ix.cpu().numpy(), va.cpu().numpy(), gr.cpu().numpy(), sel.cpu().numpy(),
                                   deg=None if div is None else deg.cpu().numpy())
            got = gs.cpu().numpy()
            okb = np.allclose(got, exp, rtol=2e-5, atol=2e-5 * max(1.0, float(np.abs(exp).max())))
            print("      bwd: max err %.3g (max |exp| %.3g) %s" % (np.abs(got - exp).max() if got.size else 0.0, np.abs(exp).max() if exp.size else 0.0, "OK" if okb else "MISMATCH"), flush=True)
            assert okb


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shapes", default="reddit,yelp,flickr,proteins,products,reddit_pl")
    ap.add_argument("--ks", default="")
    ap.add_argument("--no-small", action="store_true")
    ap.add_argument("--l2fetch", type=int, default=0, help="cudaLimitMaxL2FetchGranularity (bytes) to set first")
    args = ap.parse_args()
    dev = torch.device("cuda")
    torch.zeros(1, device=dev)
    if args.l2fetch:
        rt = ctypes.CDLL("libcudart.so.12")
        got = ctypes.c_size_t(0)
        rt.cudaDeviceGetLimit(ctypes.byref(got), 5)
        st = rt.cudaDeviceSetLimit(5, ctypes.c_size_t(args.l2fetch))
        now = ctypes.c_size_t(0)
        rt.cudaDeviceGetLimit(ctypes.byref(now), 5)
        print("cudaLimitMaxL2FetchGranularity: was %d, set %d -> status %d, now %d" % (got.value, args.l2fetch, st, now.value), flush=True)
    if not args.no_small:
        small_checks()
    for shape in args.shapes.split(","):
        kind = "uniform"
        name = shape
        if shape.endswith("_pl"):
            name, kind = shape[:-3], "powerlaw"
        n, e = SHAPES[name]
        ks = [int(v) for v in args.ks.split(",")] if args.ks else ([8, 16, 32, 64] if name == "reddit" and kind == "uniform" else [64] if name == "proteins" else [32])
        g = synth_graph(n, e, seed=123, kind=kind, device=dev)
        ip, ix, va = g["indptr"], g["indices"], g["values"]
        rb, re_ = ip[:-1].contiguous(), ip[1:].contiguous()
        x = torch.rand(n, 256, device=dev, generator=torch.Generator(device=dev).manual_seed(123))
        t_plan = time_ms(lambda: build_plan(rb, re_), 1, 3)[0]
        plan = build_plan(rb, re_)
        hdr = plan[:64].view(torch.int32).tolist()
        deg = torch.clamp((ip[1:] - ip[:-1]).float(), min=1.0)
        for k in ks:
            r_old = K.topk_cbsr(x, k, order=K.ORDER_BANKED)
            r_col = K.topk_cbsr(x, k, order=K.ORDER_COLUMN_ASC)
            nv, ns = slot_order(r_col["values"], r_col["sel"])
            out_old = torch.empty(n, 256, device=dev)
            out_new = torch.empty(n, 256, device=dev)
            K.spgemm_forward_csr(rb, re_, ix, va, r_old["values"], r_old["sel"], out=out_old, row_div=deg)
            fwd_planned(plan, ix, va, nv, ns, n, out=out_new, row_div=deg)
            diff = (out_old - out_new).abs().max().item()
            rel = diff / out_old.abs().max().item()
            t_old = time_ms(lambda: K.spgemm_forward_csr(rb, re_, ix, va, r_old["values"], r_old["sel"], out=out_old))
            t_new = time_ms(lambda: fwd_planned(plan, ix, va, nv, ns, n, out=out_new))
            t_new_oldorder = time_ms(lambda: fwd_planned(plan, ix, va, r_old["values"], r_old["sel"], n, out=out_new))
            grad = torch.rand(n, 256, device=dev)
            gs_old = torch.empty(n, k, device=dev)
            gs_new = torch.empty(n, k, device=dev)
            K.sspmm_backward_csr(rb, re_, ix, va, grad, r_old["sel"], out=gs_old, row_div=deg)
            bwd_planned(plan, ix, va, grad, ns, n, row_div=deg, out=gs_new)
            d_old, d_new = K.cbsr_scatter(gs_old, r_old["sel"]), K.cbsr_scatter(gs_new, ns)
            bdiff = (d_old - d_new).abs().max().item() / d_old.abs().max().item()
            del d_old, d_new
            tb_old = time_ms(lambda: K.sspmm_backward_csr(rb, re_, ix, va, grad, r_old["sel"], out=gs_old))
            tb_new = time_ms(lambda: bwd_planned(plan, ix, va, grad, ns, n, out=gs_new))
            tb_new_oldorder = time_ms(lambda: bwd_planned(plan, ix, va, grad, r_old["sel"], n, out=gs_new))
            tb_old_neworder = time_ms(lambda: K.sspmm_backward_csr(rb, re_, ix, va, grad, ns, out=gs_old))
            r_colo = r_col["sel"]
            tb_old_colorder = time_ms(lambda: K.sspmm_backward_csr(rb, re_, ix, va, grad, r_colo, out=gs_old))
            print("%s/%s k=%d BWD: v1 %.3f ms | slots %.3f ms | slots on the v1 order %.3f | v1 on the slot order %.3f | v1 on column order %.3f | max rel diff %.2e" % (
                name, kind, k, tb_old[0], tb_new[0], tb_new_oldorder[0], tb_old_neworder[0], tb_old_colorder[0], bdiff), flush=True)
            print("%s/%s k=%d: v1 %.3f ms (min %.3f) | slots %.3f ms (min %.3f) | slots on the v1 order %.3f | max rel diff %.2e | plan %.3f ms nA=%d nB=%d nC=%d nD=%d" % (
                name, kind, k, t_old[0], t_old[1], t_new[0], t_new[1], t_new_oldorder[0], rel, t_plan, hdr[2], hdr[3], hdr[4], hdr[5]), flush=True)
        del g, ip, ix, va, x
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
