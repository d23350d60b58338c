"""GPU: the row plan (csrc/plan.cu) and the work-item modes of the slot-parallel forward against the oracle."""
import numpy as np
import pytest
import torch

import oracle
from helpers import assert_close
from synth_graphs import synth_graph

pytestmark = pytest.mark.gpu


def _plan_arrays(kern, plan):
    n = plan.n_rows
    ints = plan.buf.view(torch.int32)
    n_pad = (n + 31) // 32 * 32 + 32
    rows = ints[16:16 + n].cpu().numpy()
    beg = ints[16 + n_pad:16 + n_pad + n].cpu().numpy()
    end = ints[16 + 2 * n_pad:16 + 2 * n_pad + n].cpu().numpy()
    return rows, beg, end


@pytest.mark.parametrize("n,e,kind", [(1000, 5000, "uniform"), (5000, 400000, "powerlaw"), (300, 300 * 5000, "uniform"),
                                      (4096, 4096 * 200, "uniform"), (7, 0, "uniform")])
def test_plan_is_a_stable_degree_sort_of_the_rows(n, e, kind):
    import maxk_cuda_kernels as kern
    g = synth_graph(n, e, seed=3, kind=kind) if e else {"indptr": torch.zeros(n + 1, dtype=torch.int32)}
    ip = g["indptr"].cuda()
    plan = kern.build_plan(ip[:-1], ip[1:])
    rows, beg, end = _plan_arrays(kern, plan)
    deg = np.diff(ip.cpu().numpy())
    assert sorted(rows.tolist()) == list(range(n))                       # a permutation of the rows
    assert np.array_equal(beg, ip.cpu().numpy()[rows]) and np.array_equal(end - beg, deg[rows])

    def bucket(d):                                                       # plan_key of csrc/plan.cu, descending
        if d <= 0:
            return 0
        msb = int(d).bit_length() - 1
        return d if msb < 3 else 8 * (msb - 2) + ((d >> (msb - 3)) & 7)
    b = np.array([bucket(int(d)) for d in deg[rows]])
    assert (np.diff(b) <= 0).all()                                       # longest buckets first
    same = np.diff(b) == 0
    assert (np.diff(rows)[same] > 0).all()                               # stable inside a bucket
    h = plan.header()
    assert h["long_rows"] == int((deg >= 4096).sum())
    covered = h["long_rows"] + h["last_wave_rows"]
    assert h["items"] == h["long_rows"] + h["groups"] + h["last_wave_rows"] + h["tail_groups"]
    assert 8 * (h["groups"] + h["tail_groups"]) >= n - covered >= 8 * (h["groups"] + h["tail_groups"]) - 14


@pytest.mark.parametrize("k", [8, 16, 32, 64, 96, 128, 19, 1, 200])
@pytest.mark.parametrize("n,e,kind", [(997, 30000, "powerlaw"), (64, 64 * 4500, "uniform"), (3000, 3000 * 140, "uniform")])
def test_forward_item_modes_match_the_oracle(k, n, e, kind):
    """separate groups (short rows), shared long rows (>= 4096 edges), shared last-wave rows (a regular graph of
    degree >= 128), for the vectorised k and the scalar-load path."""
    import maxk_cuda_kernels as kern
    g = synth_graph(n, e, seed=k, kind=kind)
    ip, ix, va = g["indptr"].cuda(), g["indices"].cuda(), g["values"].cuda()
    x = torch.randn(n, 256, generator=torch.Generator().manual_seed(k))
    vals, cols = oracle.topk(x.numpy(), k, 2)
    sel = cols.astype(np.uint8)
    deg = np.maximum(np.diff(g["indptr"].numpy()), 1).astype(np.float32)
    plan = kern.build_plan(ip[:-1], ip[1:])
    out = kern.spgemm_forward_csr(ip[:-1], ip[1:], ix, va, torch.from_numpy(vals).cuda(), torch.from_numpy(sel).cuda(),
                                  row_div=torch.from_numpy(deg).cuda(), plan=plan)
    exp = oracle.spgemm_fwd(g["indptr"].numpy(), g["indices"].numpy(), g["values"].numpy(), vals, sel, deg=deg)
    assert_close(out, exp, "forward k=%d" % k, rtol=2e-5)
    h = plan.header()
    if n == 64:
        assert h["long_rows"] == 64
    if n == 3000:
        assert h["last_wave_rows"] == 3000


def test_plan_of_another_graph_is_rejected():
    import maxk_cuda_kernels as kern
    g = synth_graph(100, 500, seed=1)
    ip = g["indptr"].cuda()
    plan = kern.build_plan(ip[:-1], ip[1:])
    with pytest.raises(RuntimeError):
        kern.spgemm_forward_csr(ip[:-2], ip[1:-1], g["indices"].cuda(), g["values"].cuda(), torch.zeros(100, 32).cuda(),
                                torch.zeros(100, 32, dtype=torch.uint8).cuda(), plan=plan)


@pytest.mark.parametrize("k", [96, 128])
@pytest.mark.parametrize("n,e,kind", [(997, 30000, "powerlaw"), (64, 64 * 4500, "uniform")])
def test_backward_wide_rows_match_the_oracle(k, n, e, kind):
    """k = 96 / 128: the backward runs the k = 32 lane layout over 3 / 4 chunks of every CBSR row."""
    import maxk_cuda_kernels as kern
    g = synth_graph(n, e, seed=k, kind=kind)
    ip, ix, va = g["indptr"].cuda(), g["indices"].cuda(), g["values"].cuda()
    gen = torch.Generator().manual_seed(k)
    x, grad = torch.randn(n, 256, generator=gen), torch.rand(n, 256, generator=gen)
    _, cols = oracle.topk(x.numpy(), k, 2)
    sel = cols.astype(np.uint8)
    deg = np.maximum(np.diff(g["indptr"].numpy()), 1).astype(np.float32)
    gs = kern.sspmm_backward_csr(ip[:-1], ip[1:], ix, va, grad.cuda(), torch.from_numpy(sel).cuda(),
                                 row_div=torch.from_numpy(deg).cuda())
    exp = oracle.sspmm_bwd(g["indptr"].numpy(), g["indices"].numpy(), g["values"].numpy(), grad.numpy(), sel, deg=deg)
    assert_close(gs, exp, "backward k=%d" % k, rtol=2e-5)
