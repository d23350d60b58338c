"""One top-k launch on the Reddit-shape feature matrix (ncu target)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "spgemm-prunning_b200"))
import maxk_cuda_kernels as K  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 232965
k = int(sys.argv[2]) if len(sys.argv) > 2 else 32
dist = sys.argv[3] if len(sys.argv) > 3 else "uniform"
torch.manual_seed(123)
x = torch.rand(n, 256, device="cuda") if dist == "uniform" else torch.randn(n, 256, device="cuda")
for _ in range(3):
    r = K.topk_cbsr(x, k, order=2)
torch.cuda.synchronize()
print("ok", r["values"].shape)
