"""GPU: feature widths above 256 -- uint16 selectors (csrc/wide.cu) against the oracle."""
import numpy as np
import pytest
import torch

import oracle
from helpers import assert_close
from synth_graphs import synth_graph

pytestmark = pytest.mark.gpu


def _sel16(t):
    return (t.to(torch.int32) & 0xffff).cpu().numpy().astype(np.uint16)


@pytest.mark.parametrize("dim,k", [(384, 32), (384, 19), (257, 8), (512, 64), (1000, 32), (1024, 256), (100, 16)])
def test_topk16_exact(dim, k):
    import maxk_cuda_kernels as kern
    gen = torch.Generator().manual_seed(dim + k)
    x = torch.randn(300, dim, generator=gen)
    x[0, : dim // 2] = 1.5                                   # ties across rank k: lowest columns win
    x[1] = 0.0
    x[2, 5] = float("nan")
    x[3, ::7] = float("inf")
    x[4, ::3] = -0.0
    r = kern.topk_cbsr16(x.cuda(), k, want_masked=True)
    vals, cols = oracle.topk(x.numpy(), k, 1)              # column-ascending
    assert np.array_equal(_sel16(r["sel16"]), cols.astype(np.uint16))
    assert np.array_equal(r["values"].cpu().numpy().view(np.uint32), vals.view(np.uint32))
    exp_masked = oracle.maxk_act_fwd(x.numpy(), cols)
    got = r["masked"].cpu().numpy()
    assert np.array_equal(np.isnan(got), np.isnan(exp_masked)) and np.array_equal(np.nan_to_num(got), np.nan_to_num(exp_masked))


@pytest.mark.parametrize("dim,k", [(384, 32), (300, 19), (1000, 64), (512, 128)])
@pytest.mark.parametrize("n,e,kind", [(600, 20000, "powerlaw"), (64, 64 * 4200, "uniform")])
def test_wide_forward_backward_vs_oracle(dim, k, n, e, kind):
    import maxk_cuda_kernels as kern
    g = synth_graph(n, e, seed=dim + k, kind=kind)
    ip, ix, va = g["indptr"].cuda(), g["indices"].cuda(), g["values"].cuda()
    gen = torch.Generator().manual_seed(k)
    x, grad = torch.randn(n, dim, generator=gen), torch.rand(n, dim, generator=gen)
    deg = torch.clamp((g["indptr"][1:] - g["indptr"][:-1]).float(), min=1)
    r = kern.topk_cbsr16(x.cuda(), k)
    sel16 = _sel16(r["sel16"])
    vals = r["values"].cpu().numpy()
    ipn, ixn, van = g["indptr"].numpy(), g["indices"].numpy(), g["values"].numpy()
    out = kern.spgemm_forward16_csr(ip[:-1], ip[1:], ix, va, r["values"], r["sel16"], dim, row_div=deg.cuda())
    assert_close(out, oracle.spgemm_fwd16(ipn, ixn, van, vals, sel16, dim, deg=deg.numpy()), "wide forward", rtol=2e-5)
    gs = kern.sspmm_backward16_csr(ip[:-1], ip[1:], ix, va, grad.cuda(), r["sel16"], row_div=deg.cuda())
    assert_close(gs, oracle.sspmm_bwd16(ipn, ixn, van, grad.numpy(), sel16, deg=deg.numpy()), "wide backward", rtol=2e-5)


def test_wide_operator_autograd_hidden_384():
    """The reference's Yelp script trains with hidden 384 (scripts_train/yelp_maxk.sh:16), which its uint8 selectors
    cannot address: the wide operator against torch.topk + torch.sparse on the same inputs."""
    import maxk_cuda_kernels as kern
    n, dim, k = 800, 384, 32
    g = synth_graph(n, 16000, seed=9, kind="powerlaw")
    ip, ix, va = g["indptr"].cuda(), g["indices"].cuda(), g["values"].cuda()
    deg = torch.clamp((ip[1:] - ip[:-1]).float(), min=1)
    x = torch.randn(n, dim, device="cuda", requires_grad=True)
    up = torch.rand(n, dim, device="cuda")
    out = kern.WideMaxKSpGEMMFunction.apply(ip, ix, va, x, k, deg, deg)
    out.backward(up)
    got_grad = x.grad.clone()
    x.grad = None
    a = torch.sparse_csr_tensor(ip.long(), ix.long(), va, size=(n, n))
    v, i = torch.topk(x, k, dim=1)
    xs = torch.zeros_like(x).scatter(1, i, v)
    ref = (a @ xs) / deg.unsqueeze(-1)
    ref.backward(up)
    torch.testing.assert_close(out, ref, rtol=1e-5, atol=1e-6)
    # d/dx of A @ (x * mask) / deg:  A^T (up / deg) at the selected positions
    torch.testing.assert_close(got_grad, x.grad, rtol=2e-5, atol=1e-6)
