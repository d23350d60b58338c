"""cpu_baseline.py -- the reference's CPU path for this hot path, timed on host cores.

TEST/BENCH INFRASTRUCTURE ONLY (see maxk_oracle.c header).

BASELINE.json names "DGL gspmm plus torch.topk" as the reference's CPU path
(maxk_models_integrated.py:32 MaxK top-k, :314-320 update_all(copy_u, mean) -> DGL gspmm).
DGL is not installed and cannot be (no network), so this is a PORT: torch.topk is the same
call; gspmm is restated as torch.sparse_csr @ dense (MKL, all host threads):
    fwd:  v, i = torch.topk(x, k, 1); xs = zeros.scatter_(1, i, v); out = A_csr @ xs
    bwd:  gd = A_csr^T @ g (transpose pre-built, untimed);      gs = gd.gather(1, i)
"""
import os
import time

import torch


def make_problem(graph, k, dim=256, seed=123):
    """Host tensors for one fwd+bwd aggregation on `graph` (dict from synth_graphs, CPU tensors)."""
    n = graph["v_num"]
    a = torch.sparse_csr_tensor(graph["indptr"].to(torch.int64).cpu(), graph["indices"].to(torch.int64).cpu(),
                                graph["values"].cpu(), size=(n, n))
    at = a.to_sparse_coo().t().coalesce().to_sparse_csr()
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(n, dim, generator=g)
    grad = torch.rand(n, dim, generator=g)
    return {"a": a, "at": at, "x": x, "grad": grad, "k": k}


def run_layer(p):
    """One top-k + forward aggregation + backward aggregation on the CPU. Returns (out, gs)."""
    v, i = torch.topk(p["x"], p["k"], dim=1)
    xs = torch.zeros_like(p["x"]).scatter_(1, i, v)
    out = p["a"] @ xs
    gd = p["at"] @ p["grad"]
    gs = gd.gather(1, i)
    return out, gs


def time_layer(p, steps=1, warmup=0, threads=None):
    threads = threads or os.cpu_count() or 1
    torch.set_num_threads(threads)
    for _ in range(warmup):
        run_layer(p)
    t0 = time.perf_counter()
    for _ in range(steps):
        run_layer(p)
    return (time.perf_counter() - t0) / max(steps, 1), threads
