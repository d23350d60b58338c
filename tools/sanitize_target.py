"""Small run of every kernel family for compute-sanitizer (one tool per gpurun call)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "spgemm-prunning_b200"),):
    sys.path.insert(0, p)
import maxk_cuda_kernels as K  # noqa: E402
from synth_graphs import synth_graph  # noqa: E402

g = synth_graph(600, 30000, seed=1, kind="powerlaw", device="cuda")
deg = np.zeros(600, np.int64)
ip, ix, va = g["indptr"], g["indices"], g["values"]
x = torch.randn(600, 256, device="cuda")
grad = torch.rand(600, 256, device="cuda")
for k in (8, 16, 32, 64, 19):
    for order in (0, 1, 2):
        r = K.topk_cbsr(x, k, order=order, want_masked=True, want_i32=True, want_i64=True)
    out = K.spgemm_forward_csr(ip[:-1], ip[1:], ix, va, r["values"], r["sel"])
    gs = K.sspmm_backward_csr(ip[:-1], ip[1:], ix, va, grad, r["sel"])
    K.cbsr_scatter(gs, r["sel"])
    K.mask_apply(grad, r["sel"], gs)
# one long row (CTA path)
n = 64
ptr = torch.zeros(n + 1, dtype=torch.int32, device="cuda")
ptr[5:] = 5000
idx = torch.randint(0, n, (5000,), device="cuda", dtype=torch.int32)
val = torch.rand(5000, device="cuda")
r = K.topk_cbsr(torch.randn(n, 256, device="cuda"), 32)
K.spgemm_forward_csr(ptr[:-1], ptr[1:], idx, val, r["values"], r["sel"])
K.sspmm_backward_csr(ptr[:-1], ptr[1:], idx, val, torch.rand(n, 256, device="cuda"), r["sel"])
w4, nw = K.build_warp4(ip, 64)
K.spmm_maxk_forward(w4, ix, va, K.topk_cbsr(x, 32)["values"], K.topk_cbsr(x, 32)["sel"], nw, 32)
torch.cuda.synchronize()
print("sanitize target done")
