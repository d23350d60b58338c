"""cpu_baseline.py -- the reference's CPU path for this hot path, timed on host cores.

TEST/BENCH INFRASTRUCTURE ONLY (see maxk_oracle.c header): used by bench.py's cpu_baseline leg and
by `bench.py --impl reference`, never by the product.

BASELINE.json names "DGL gspmm plus torch.topk" as the reference's CPU path
(maxk_models_integrated.py:32 MaxK top-k, :314-320 update_all(copy_u, mean) -> DGL gspmm).
DGL is not installed and cannot be (no network), so this is a PORT (kind = "port"): torch.topk is
the very same call; gspmm is restated as torch.sparse_csr @ dense (MKL, all host threads):
    fwd:  v, i = torch.topk(x, k, 1); xs = zeros.scatter_(1, i, v); out = A_s @ xs
    bwd:  gd = A_s^T @ g_s   (transpose pre-built, untimed);            gs = gd.gather(1, i)
A "sample" is the first n_rows rows of an identically distributed graph over all n_total nodes
(uniform kind: every row has avg_deg iid uniform neighbours), so one step does the full top-k
(every node can be a neighbour) and n_rows/n_total of the aggregation work.
"""
import os
import time

import warnings

import torch

warnings.filterwarnings("ignore", message=".*[Ss]parse.*")


def algorithmic_bytes(n_total, n_rows, n_edges, k, dim=256):
    """SURVEY.md 8(d) formulas for a row slab: top-k over all nodes, fwd + bwd over n_rows rows."""
    b_topk = n_total * dim * 4 + n_total * k * 5
    b_fwd = (n_rows + 1) * 4 + n_edges * 8 + n_total * k * 5 + n_rows * dim * 4
    b_bwd = (n_rows + 1) * 4 + n_edges * 8 + n_rows * dim * 4 + n_total * k * 5
    return b_topk + b_fwd + b_bwd


def sample_problem(n_total, avg_deg, n_rows, k, dim=256, seed=123):
    g = torch.Generator().manual_seed(seed)
    n_rows = max(1, min(n_rows, n_total))
    e = n_rows * avg_deg
    cols, _ = torch.sort(torch.randint(0, n_total, (n_rows, avg_deg), generator=g), dim=1)
    indptr = torch.arange(0, e + 1, avg_deg, dtype=torch.int64)
    vals = torch.rand(e, generator=g)
    a = torch.sparse_csr_tensor(indptr, cols.reshape(-1), vals, size=(n_rows, n_total))
    at = a.to_sparse_coo().t().coalesce().to_sparse_csr()
    x = torch.rand(n_total, dim, generator=g)
    grad = torch.rand(n_rows, dim, generator=g)
    return {"a": a, "at": at, "x": x, "grad": grad, "k": k, "n_total": n_total, "n_rows": n_rows, "n_edges": e,
            "bytes": algorithmic_bytes(n_total, n_rows, e, k, dim)}


def run_layer(p):
    v, i = torch.topk(p["x"], p["k"], dim=1)
    xs = torch.zeros_like(p["x"]).scatter_(1, i, v)
    out = p["a"] @ xs
    gd = p["at"] @ p["grad"]
    gs = gd.gather(1, i)
    return out, gs


def time_layer(p, steps=1, warmup=0, threads=None):
    """Returns (seconds per step, threads used)."""
    threads = threads or os.cpu_count() or 1
    torch.set_num_threads(threads)
    for _ in range(warmup):
        run_layer(p)
    t0 = time.perf_counter()
    for _ in range(steps):
        run_layer(p)
    return (time.perf_counter() - t0) / max(steps, 1), threads
