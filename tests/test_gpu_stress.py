"""GPU: randomized parity sweep against the oracle (ragged / empty / long rows, odd sizes, every k path)."""
import numpy as np
import pytest
import torch

import oracle

pytestmark = pytest.mark.gpu


def _case(seed):
    rng = np.random.default_rng(seed)
    n = int(rng.choice([1, 2, 7, 31, 33, 257, 1000, 4097]))
    n_src = int(rng.choice([n, max(1, n // 2), n + 13]))             # rectangular: sources != rows (sharded slabs)
    k = int(rng.choice([1, 2, 4, 8, 12, 16, 19, 32, 33, 64, 100, 256]))
    style = rng.choice(["sparse", "dense", "skew", "empty"])
    if style == "sparse":
        deg = rng.integers(0, 6, n)
    elif style == "dense":
        deg = rng.integers(0, 400, n)
    elif style == "skew":
        deg = rng.integers(0, 8, n)
        deg[rng.integers(0, n, max(1, n // 50))] = rng.integers(4000, 9000)   # beyond kLongRow: CTA path
    else:
        deg = np.zeros(n, np.int64)
    deg[rng.random(n) < 0.15] = 0
    indptr = np.zeros(n + 1, np.int32)
    indptr[1:] = np.cumsum(deg)
    e = int(indptr[-1])
    indices = np.sort(rng.integers(0, n_src, e)).astype(np.int32) if rng.random() < 0.3 else rng.integers(0, n_src, e).astype(np.int32)
    values = rng.standard_normal(e).astype(np.float32)
    x = rng.standard_normal((n_src, 256)).astype(np.float32)
    if rng.random() < 0.3:
        x = np.round(x * 4) / 4                                        # ties
    grad = rng.standard_normal((n, 256)).astype(np.float32)
    order = int(rng.choice([0, 1, 2]))
    use_div = rng.random() < 0.5
    div = (rng.random(n).astype(np.float32) * 9 + 1) if use_div else None
    return n, n_src, k, indptr, indices, values, x, grad, order, div


@pytest.mark.parametrize("seed", range(48))
def test_random_case(seed):
    import maxk_cuda_kernels as kern
    n, n_src, k, indptr, indices, values, x, grad, order, div = _case(seed)
    t = lambda a: torch.from_numpy(a).cuda()
    ev, ec = oracle.topk(x, k, order)
    r = kern.topk_cbsr(t(x), k, order=order, want_masked=True)
    assert np.array_equal(r["sel"].cpu().numpy(), ec.astype(np.uint8)), "top-k selectors"
    assert np.array_equal(r["values"].cpu().numpy(), ev), "top-k values"
    assert np.array_equal(r["masked"].cpu().numpy(), oracle.maxk_act_fwd(x, ec)), "masked row"
    sel = ec.astype(np.uint8)
    cip, cix, cva = t(indptr), t(indices), t(values)
    d = t(div) if div is not None else None
    # signed data cancels, so the 1e-5 relative bound is taken against the sum of |terms| (the oracle run on
    # absolute values), which is what bounds the rounding error of any summation order
    out = kern.spgemm_forward_csr(cip[:-1], cip[1:], cix, cva, r["values"], r["sel"], row_div=d).cpu().numpy()
    exp = oracle.spgemm_fwd(indptr, indices, values, ev, sel, deg=div)
    mag = oracle.spgemm_fwd(indptr, indices, np.abs(values), np.abs(ev), sel, deg=div)
    assert out.shape == exp.shape and (np.abs(out - exp) <= 1e-5 * mag + 1e-6).all(), "forward"
    gs = kern.sspmm_backward_csr(cip[:-1], cip[1:], cix, cva, t(grad), r["sel"], row_div=d).cpu().numpy()
    assert gs.shape == (n_src, k)
    expb = oracle.sspmm_bwd(indptr, indices, values, grad, sel, deg=div)
    magb = oracle.sspmm_bwd(indptr, indices, np.abs(values), np.abs(grad), sel, deg=div)
    assert (np.abs(gs - expb) <= 1e-5 * magb + 1e-6).all(), "backward"
