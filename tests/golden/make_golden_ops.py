"""Generates tests/golden/ref_ops.npz: the REFERENCE's own autograd operators (their Python glue: where the
degree divisions happen, what is saved, what the backward returns) executed on the CPU.

Run in the build container only (needs /root/reference; never at test time):
    python tests/golden/make_golden_ops.py

The reference's operator files import a compiled extension `maxk_cuda_kernels` that cannot be built here
(SURVEY 8c).  A stand-in module with the same two entry points is put into sys.modules; it evaluates the
kernels' mathematical contract (SURVEY a-4 / a-6, the formulas the golden CUDA vectors of ref_cuda.npz pin)
in float64 with scipy, driven by the warp4 quads exactly like the kernels are.  Everything else that runs
is the reference's unmodified code, loaded from where it lies:
  * spgemmfunction_v3.py   MaxKSpGEMMFunction (CSR forward / CSC + CSC-quads backward, :22-156)
  * spgemmfunction_v4      MaxKSpGEMMFunction (pre-computed top-k, undirected graphs, :19-101)
  * spgemmfunction.py      OptimizedMaxKSpGEMMFunction (:18-108)
Inputs are seeded and tie-free.  Nothing of the reference is copied into this repository.
"""
import contextlib
import importlib.machinery
import io
import os
import sys
import types

import numpy as np
import scipy.sparse as sp
import torch

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "spgemm-prunning_b200"))
from graph_loader import csr_to_csc  # noqa: E402   (host-side helpers only, NOT our maxk_cuda_kernels)
from synth_graphs import symmetrize, synth_graph  # noqa: E402

sys.path.remove(os.path.join(ROOT, "spgemm-prunning_b200"))
sys.modules.pop("maxk_cuda_kernels", None)


def _matrix(warp4, indices, values, n):
    """The CSR matrix the quads describe: segment (row, loc, len) covers edges [loc, loc+len) of `row`."""
    q = warp4.numpy().reshape(-1, 4)
    rows = np.repeat(q[:, 0], q[:, 2])
    pos = np.concatenate([np.arange(lo, lo + ln) for lo, ln in zip(q[:, 1], q[:, 2])]) if len(q) else np.zeros(0, np.int64)
    return sp.csr_matrix((values.numpy().astype(np.float64)[pos], (rows, indices.numpy()[pos])), shape=(n, n))


def spmm_maxk_forward(warp4, indices, values, data, selector, num_warps, k):
    n = data.shape[0]
    dense = np.zeros((n, 256))
    np.put_along_axis(dense, selector.numpy().astype(np.int64), data.detach().numpy().astype(np.float64), axis=1)
    return torch.from_numpy((_matrix(warp4, indices, values, n) @ dense).astype(np.float32))


def spmm_maxk_backward(warp4, indices, values, grad, selector, num_warps, k):
    n = selector.shape[0]
    full = _matrix(warp4, indices, values, n).T @ grad.detach().numpy().astype(np.float64)
    return torch.from_numpy(np.take_along_axis(full, selector.numpy().astype(np.int64), axis=1).astype(np.float32))


def _quads(indptr):
    out = []
    ip = indptr.numpy()
    for r in range(len(ip) - 1):                      # kernels/generate_meta.py:30-48
        loc, deg = int(ip[r]), int(ip[r + 1] - ip[r])
        while deg > 0:
            ln = min(64, deg)
            out.append((r, loc, ln, 0))
            loc, deg = loc + ln, deg - ln
    return torch.tensor(out, dtype=torch.int32).reshape(-1)


def _load(name, filename):
    return importlib.machinery.SourceFileLoader(name, os.path.join(REF, filename)).load_module()


def main():
    stand_in = types.ModuleType("maxk_cuda_kernels")
    stand_in.spmm_maxk_forward = spmm_maxk_forward
    stand_in.spmm_maxk_backward = spmm_maxk_backward
    sys.modules["maxk_cuda_kernels"] = stand_in
    out = {}
    gen = torch.Generator().manual_seed(11)
    k, d = 32, 256

    # ---- v3: directed graph, CSR forward, CSC arrays + CSC quads backward ---------------------------------
    g = synth_graph(150, 2500, seed=9, kind="powerlaw")
    ip, ix, va = g["indptr"], g["indices"], g["values"]
    t_ptr, t_idx, t_val = csr_to_csc(ip, ix, va)
    in_deg = torch.clamp((ip[1:] - ip[:-1]).float(), min=1)
    out_deg = torch.clamp((t_ptr[1:] - t_ptr[:-1]).float(), min=1)
    x = torch.randn(150, d, generator=gen)
    up = torch.rand(150, d, generator=gen)
    w_csr, w_csc = _quads(ip), _quads(t_ptr)
    with contextlib.redirect_stdout(io.StringIO()):
        v3 = _load("ref_spgemmfunction_v3", "spgemmfunction_v3.py")
        assert v3.MAXK_KERNELS_AVAILABLE
        xi = x.clone().requires_grad_(True)
        y = v3.maxk_spgemm(ix, va, xi, k, w_csr, w_csr.numel() // 4, ip, in_deg, out_deg, t_idx, t_val,
                           w_csc, w_csc.numel() // 4)
        y.backward(up)
    out.update(v3_indptr=ip.numpy(), v3_indices=ix.numpy(), v3_values=va.numpy(), v3_x=x.numpy(), v3_up=up.numpy(),
               v3_out=y.detach().numpy(), v3_grad_input=xi.grad.numpy())

    # ---- v4 and the "optimized" operator: undirected graph, pre-computed top-k ---------------------------
    gu = symmetrize(synth_graph(180, 1500, seed=4))
    ip, ix, va = gu["indptr"], gu["indices"], gu["values"]
    deg = torch.clamp((ip[1:] - ip[:-1]).float(), min=1)
    x = torch.randn(180, d, generator=gen)
    up = torch.rand(180, d, generator=gen)
    tv0, ti = torch.topk(x, k, dim=1)
    w = _quads(ip)
    with contextlib.redirect_stdout(io.StringIO()):
        v4 = _load("ref_spgemmfunction_v4", "spgemmfunction_v4")
        tv = tv0.clone().requires_grad_(True)
        y4 = v4.maxk_spgemm(ix, va, tv, ti, w, w.numel() // 4, ip, deg)
        y4.backward(up)
        g4 = tv.grad.clone()
        opt = _load("ref_spgemmfunction_opt", "spgemmfunction.py")
        tv = tv0.clone().requires_grad_(True)
        t_ptr, t_idx, t_val = csr_to_csc(ip, ix, va)
        yo = opt.optimized_maxk_spgemm(ix, va, tv, ti, w, w.numel() // 4, ip, deg, deg, t_idx, t_val)
        yo.backward(up)
        go = tv.grad.clone()
    out.update(u_indptr=ip.numpy(), u_indices=ix.numpy(), u_values=va.numpy(), u_x=x.numpy(), u_up=up.numpy(),
               u_topk_values=tv0.numpy(), u_topk_indices=ti.numpy(),
               v4_out=y4.detach().numpy(), v4_grad_topk_values=g4.numpy(),
               opt_out=yo.detach().numpy(), opt_grad_topk_values=go.numpy())
    np.savez_compressed(os.path.join(HERE, "ref_ops.npz"), **out)
    print("wrote ref_ops.npz:", {key: v.shape for key, v in out.items()})


if __name__ == "__main__":
    main()
