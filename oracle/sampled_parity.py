"""Sampled-row parity of the CUDA path against the C oracle at FULL problem sizes (test infrastructure).

The oracle walks one row at a time on the CPU, so full-size graphs (114.6 M edges) are checked on a random
sample of rows: the sub-problem those rows span (their edges, the CBSR rows / gradient rows they touch) is
cut out on the GPU, copied to the host and recomputed by oracle/maxk_oracle.c (which follows
kernels/spmm_maxk.cu:62-105 and kernels/spmm_maxk_backward.cu:52-113 of the reference).  Used by tests/ and,
outside the timed region, by bench.py's checker leg; never by the product path.

Tolerance is north_star's: top-k index sets bit-exact, fp32 within rtol 1e-5 / atol 1e-6.  `violation` below is
max |got - exp| / (atol + rtol |exp|): parity holds iff it is <= 1.
"""
import numpy as np
import torch

import oracle

RTOL, ATOL = 1e-5, 1e-6


def _violation(got, exp):
    got = np.asarray(got, dtype=np.float64)
    exp = np.asarray(exp, dtype=np.float64)
    if got.size == 0:
        return 0.0, 0.0
    d = np.abs(got - exp)
    return float((d / (ATOL + RTOL * np.abs(exp))).max()), float(d.max() / max(np.abs(exp).max(), 1e-30))


def sample_ids(n, count, seed, device):
    g = torch.Generator(device="cpu").manual_seed(seed)
    count = min(count, n)
    return torch.randperm(n, generator=g)[:count].sort().values.to(device)


def check_topk(x_rows, vals_rows, sel_rows, k):
    """x_rows [s, D], vals_rows / sel_rows [s, k] (any entry order) -> (index sets equal, values equal)."""
    x = x_rows.detach().float().cpu().numpy()
    exp_v, exp_c = oracle.topk(x, k, 1)                     # column-ascending order
    sel = sel_rows.detach().cpu().numpy().astype(np.int64)
    val = vals_rows.detach().cpu().numpy()
    order = np.argsort(sel, axis=1, kind="stable")
    got_c = np.take_along_axis(sel, order, 1)
    got_v = np.take_along_axis(val, order, 1)
    sets_equal = bool(np.array_equal(got_c, exp_c.astype(np.int64)))
    vals_equal = bool(np.array_equal(got_v.view(np.uint32), exp_v.view(np.uint32))) if sets_equal else False
    return sets_equal, vals_equal


def check_forward(indptr, indices, values, cbsr_val, cbsr_sel, rows, out_rows, row_div=None, dim=256):
    """out_rows [s, dim] = the kernel's output for CSR rows `rows` (int64 ids).  All tensors live on one device.
    Returns (violation, max_rel)."""
    rows = rows.long()
    ip = indptr.long()
    b, e = ip[rows], ip[rows + 1]
    deg = e - b
    sub_ptr = torch.zeros(rows.numel() + 1, dtype=torch.int64, device=rows.device)
    sub_ptr[1:] = torch.cumsum(deg, 0)
    total = int(sub_ptr[-1])
    # edge positions of the sampled rows
    pos = torch.arange(total, device=rows.device) - torch.repeat_interleave(sub_ptr[:-1], deg) + torch.repeat_interleave(b, deg)
    src = indices[pos].long()
    uniq, inv = torch.unique(src, return_inverse=True)
    exp = oracle.spgemm_fwd(sub_ptr.to(torch.int32).cpu().numpy(), inv.to(torch.int32).cpu().numpy(),
                            values[pos].cpu().numpy(), cbsr_val[uniq].cpu().numpy(), cbsr_sel[uniq].cpu().numpy(), dim=dim,
                            deg=None if row_div is None else row_div[rows].cpu().numpy())
    return _violation(out_rows.detach().cpu().numpy(), exp)


def check_backward(indptr, indices, values, grad, cbsr_sel, dst, gs_rows, row_div=None, max_edges=400_000):
    """gs_rows [s, k] = the kernel's sampled gradient for destination nodes `dst`; grad [n_rows, dim] is the upstream
    gradient of ALL source rows of the CSR.  Incoming edges are found on the GPU.  Returns (violation, max_rel, s_used)."""
    dst = dst.long()
    hit = torch.isin(indices, dst.to(indices.dtype))
    pos = hit.nonzero(as_tuple=False).squeeze(1)
    if pos.numel() > max_edges:                         # dense graph: keep a prefix of the sample that fits
        d_of = indices[pos].long()
        cnt = torch.bincount(torch.searchsorted(dst, d_of), minlength=dst.numel())
        keep = int((torch.cumsum(cnt, 0) <= max_edges).sum().clamp(min=1))
        dst = dst[:keep]
        gs_rows = gs_rows[:keep]
        m = torch.isin(indices[pos], dst.to(indices.dtype))
        pos = pos[m]
    src_row = torch.searchsorted(indptr.long(), pos, right=True) - 1
    uniq_rows, row_inv = torch.unique(src_row, return_inverse=True)
    # sub-CSR over the distinct source rows (edges are already grouped by row: pos is ascending)
    counts = torch.bincount(row_inv, minlength=uniq_rows.numel())
    sub_ptr = torch.zeros(uniq_rows.numel() + 1, dtype=torch.int64, device=dst.device)
    sub_ptr[1:] = torch.cumsum(counts, 0)
    local_dst = torch.searchsorted(dst, indices[pos].long())
    exp = oracle.sspmm_bwd(sub_ptr.to(torch.int32).cpu().numpy(), local_dst.to(torch.int32).cpu().numpy(),
                           values[pos].cpu().numpy(), grad[uniq_rows].float().cpu().numpy(), cbsr_sel[dst].cpu().numpy(),
                           deg=None if row_div is None else row_div[uniq_rows].cpu().numpy())
    v, r = _violation(gs_rows.detach().cpu().numpy(), exp)
    return v, r, int(dst.numel())
