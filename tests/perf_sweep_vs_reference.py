"""Developer timing sweep (not the contract bench, not collected by pytest): ours vs the reference kernels
recompiled for sm_100a (oracle/_ref).  Lives under tests/ because only test code may execute oracle/."""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "spgemm-prunning_b200"), os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)

import maxk_cuda_kernels as K  # noqa: E402
import oracle  # noqa: E402
from synth_graphs import SHAPES, synth_graph  # noqa: E402


def timeit(fn, warm=3, reps=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        b.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2], ts[0]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shape", default="reddit")
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--kind", default="uniform")
    ap.add_argument("--ks", default="8,16,32,64")
    ap.add_argument("--ref", type=int, default=1)
    ap.add_argument("--check", type=int, default=1)
    a = ap.parse_args()
    n, e = SHAPES[a.shape]
    n, e = int(n * a.scale), int(e * a.scale)
    g = synth_graph(n, e, seed=123, kind=a.kind, device="cuda")
    ip, ix, va = g["indptr"], g["indices"], g["values"]
    torch.manual_seed(123)
    x = torch.rand(n, 256, device="cuda")
    grad = torch.rand(n, 256, device="cuda")
    w4, nw = K.build_warp4(ip, 64)
    plan = K.build_plan(ip[:-1], ip[1:])            # once per graph, like the reference's metadata
    print("graph %s n=%d e=%d W=%d maxdeg=%d" % (a.shape, n, e, nw, int((ip[1:] - ip[:-1]).max())))
    for k in [int(s) for s in a.ks.split(",")]:
        r = K.topk_cbsr(x, k, order=2)
        data, sel = r["values"], r["sel"]
        r0 = K.topk_cbsr(x, k, order=0)
        t_topk = timeit(lambda: K.topk_cbsr(x, k, order=2))
        t_topk0 = timeit(lambda: K.topk_cbsr(x, k, order=0))
        t_torch = timeit(lambda: torch.topk(x, k, dim=1))
        t_fwd = timeit(lambda: K.spgemm_forward_csr(ip[:-1], ip[1:], ix, va, data, sel, plan=plan))
        t_bwd = timeit(lambda: K.sspmm_backward_csr(ip[:-1], ip[1:], ix, va, grad, sel))
        t_fwd0 = timeit(lambda: K.spgemm_forward_csr(ip[:-1], ip[1:], ix, va, r0["values"], r0["sel"], plan=plan))
        t_bwd0 = timeit(lambda: K.sspmm_backward_csr(ip[:-1], ip[1:], ix, va, grad, r0["sel"]))
        bt = n * 256 * 4 + n * k * 5
        bf = (n + 1) * 4 + e * 8 + n * k * 5 + n * 256 * 4
        bb = (n + 1) * 4 + e * 8 + n * 256 * 4 + n * k * 5
        line = "k=%d topk %.3f ms (%.0f GB/s; sorted %.3f; torch %.3f) | fwd %.3f ms (%.0f GB/s alg, gather %.0f GB/s) | bwd %.3f ms (%.0f GB/s alg)" % (
            k, t_topk[0], bt / t_topk[0] / 1e6, t_topk0[0], t_torch[0], t_fwd[0], bf / t_fwd[0] / 1e6,
            e * k * 5 / t_fwd[0] / 1e6, t_bwd[0], bb / t_bwd[0] / 1e6)
        line += " | value-order CBSR: fwd %.3f bwd %.3f" % (t_fwd0[0], t_bwd0[0])
        if a.ref and oracle.ref_cuda_available():
            t_rf = timeit(lambda: oracle.ref_cuda_forward(w4, ix, va, data, sel, nw), warm=2, reps=3)
            t_rb = timeit(lambda: oracle.ref_cuda_backward(w4, ix, va, grad, sel, nw), warm=2, reps=3)
            line += " | REF fwd %.3f bwd %.3f ms" % (t_rf[0], t_rb[0])
            if a.check:
                o = K.spgemm_forward_csr(ip[:-1], ip[1:], ix, va, data, sel, plan=plan)
                ro = oracle.ref_cuda_forward(w4, ix, va, data, sel, nw)
                gs = K.sspmm_backward_csr(ip[:-1], ip[1:], ix, va, grad, sel)
                rgs = oracle.ref_cuda_backward(w4, ix, va, grad, sel, nw)
                line += " | maxrel fwd %.2e bwd %.2e" % (
                    float(((o - ro).abs() / (ro.abs() + 1e-6)).max()), float(((gs - rgs).abs() / (rgs.abs() + 1e-6)).max()))
        print(line, flush=True)


if __name__ == "__main__":
    main()
