"""Developer tool (GPU): time the forward / backward on BASELINE shapes (median of 8, CUDA events).
    python tools/fwd_once.py [--shapes reddit,yelp,...] [--k 32 | --k 8,64]"""
import argparse, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "spgemm-prunning_b200")):
    sys.path.insert(0, p)
import torch
import maxk_cuda_kernels as K
from synth_graphs import SHAPES, synth_graph

def med(fn, warm=3, reps=8):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    for a, b in ev:
        a.record(); fn(); b.record()
    torch.cuda.synchronize()
    return sorted(a.elapsed_time(b) for a, b in ev)[reps // 2]

ap = argparse.ArgumentParser()
ap.add_argument("--shapes", default="reddit,yelp,flickr,proteins,products,reddit_pl")
ap.add_argument("--k", default="0", help="k, or a comma list of k (all run on every shape)")
args = ap.parse_args()
for shape, k_arg in ((s_, int(k_)) for s_ in args.shapes.split(",") for k_ in args.k.split(",")):
    name, kind = (shape[:-3], "powerlaw") if shape.endswith("_pl") else (shape, "uniform")
    n, e = SHAPES[name]
    k = k_arg or (64 if name == "proteins" else 32)
    g = synth_graph(n, e, seed=123, kind=kind, device="cuda")
    ip, ix, va = g["indptr"], g["indices"], g["values"]
    x = torch.rand(n, 256, device="cuda"); grad = torch.rand(n, 256, device="cuda")
    plan = K.build_plan(ip[:-1], ip[1:])
    r = K.topk_cbsr(x, k); out = torch.empty(n, 256, device="cuda"); gs = torch.empty(n, k, device="cuda")
    tt = med(lambda: K.topk_cbsr(x, k, out_values=r["values"], out_sel=r["sel"]))
    tf = med(lambda: K.spgemm_forward_csr(ip[:-1], ip[1:], ix, va, r["values"], r["sel"], out=out, plan=plan))
    tb = med(lambda: K.sspmm_backward_csr(ip[:-1], ip[1:], ix, va, grad, r["sel"], out=gs))
    print("%s/%s k=%d: topk %.3f fwd %.3f bwd %.3f ms" % (name, kind, k, tt, tf, tb), flush=True)
    del g, ip, ix, va, x, grad, out, gs, r, plan
    torch.cuda.empty_cache()
