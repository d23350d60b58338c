"""Developer tool (GPU): forward / backward over a 1/P row slab of the Reddit-shape graph on ONE GPU (what a rank of
the sharded layer runs), to separate kernel effects from exchange effects.   python tools/slab_fwd.py [P ...]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "spgemm-prunning_b200")):
    sys.path.insert(0, p)
import torch
import maxk_cuda_kernels as K
from synth_graphs import SHAPES, synth_graph
from sharded import shard_rows

def med(fn, warm=3, reps=10):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    for a, b in ev:
        a.record(); fn(); b.record()
    torch.cuda.synchronize()
    return sorted(a.elapsed_time(b) for a, b in ev)[reps // 2]

n, e = SHAPES["reddit"]
g = synth_graph(n, e, seed=123, kind="uniform", device="cuda")
x = torch.rand(n, 256, device="cuda")
r = K.topk_cbsr(x, 32)
for P in [int(a) for a in sys.argv[1:]] or [1, 2, 4, 8]:
    rows = shard_rows(g, P, 0)
    m = rows["v_num"]
    ip = rows["indptr"]
    plan = K.build_plan(ip[:-1], ip[1:])
    vals = torch.zeros(P * m, 32, device="cuda"); vals[:n] = r["values"]
    sel = torch.zeros(P * m, 32, dtype=torch.uint8, device="cuda"); sel[:n] = r["sel"]
    grad = torch.rand(m, 256, device="cuda")
    out = torch.empty(m, 256, device="cuda"); gs = torch.empty(P * m, 32, device="cuda")
    tf = med(lambda: K.spgemm_forward_csr(ip[:-1], ip[1:], rows["indices"], rows["values"], vals, sel, out=out, plan=plan))
    tb = med(lambda: K.sspmm_backward_csr(ip[:-1], ip[1:], rows["indices"], rows["values"], grad, sel, out=gs))
    tt = med(lambda: K.topk_cbsr(x[:m], 32))
    print("P=%d slab %d rows %d edges: fwd %.3f ms (x%d = %.3f) bwd %.3f ms (x%d = %.3f) topk %.3f  plan %s" % (
        P, m, rows["e_num"], tf, P, tf * P, tb, P, tb * P, tt, plan.header()), flush=True)
