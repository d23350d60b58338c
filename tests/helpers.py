"""Shared test helpers: seeded problems and oracle comparisons."""
import numpy as np
import torch

import oracle
from synth_graphs import synth_graph

RTOL, ATOL = 1e-5, 1e-6     # north star tolerance for fp32 outputs vs the fp64 oracle


def make_problem(n, e, k, dim=256, kind="uniform", seed=0, values="uniform", signed=False, order=2):
    g = synth_graph(n, e, seed=seed + 123, kind=kind, values=values)
    gen = torch.Generator().manual_seed(seed)
    x = torch.randn(n, dim, generator=gen) if signed else torch.rand(n, dim, generator=gen)
    grad = torch.rand(n, dim, generator=gen)
    vals, cols = oracle.topk(x.numpy(), k, order=order)
    return {"graph": g, "x": x, "grad": grad, "k": k, "dim": dim,
            "cbsr_val": vals, "cbsr_col": cols, "cbsr_sel": cols.astype(np.uint8)}


def graph_np(g):
    return g["indptr"].numpy(), g["indices"].numpy(), g["values"].numpy()


def graph_cuda(g):
    return g["indptr"].cuda(), g["indices"].cuda(), g["values"].cuda()


def assert_close(actual, expected, what, rtol=RTOL, atol=ATOL):
    actual = actual.detach().cpu().numpy() if isinstance(actual, torch.Tensor) else np.asarray(actual)
    expected = np.asarray(expected)
    assert actual.shape == expected.shape, "%s: shape %s vs %s" % (what, actual.shape, expected.shape)
    err = np.abs(actual.astype(np.float64) - expected.astype(np.float64))
    bound = atol + rtol * np.abs(expected.astype(np.float64))
    bad = err > bound
    assert not bad.any(), "%s: %d / %d elements outside rtol=%g atol=%g (max abs err %.3e)" % (
        what, int(bad.sum()), bad.size, rtol, atol, float(err.max()))
