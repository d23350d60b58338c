"""1-D row-sharded MaxK aggregation over the GPUs of one box (no reference counterpart --
the reference is single-GPU, all_train.py:224).

One process per GPU (torch.distributed, NCCL over NVLink).  Rank p owns the rows
[p*m, (p+1)*m) of the adjacency (m = ceil(N / P); the last rank's slab is zero-padded so every
collective has uniform counts), the same rows of the features and of the output.

forward   : local top-k -> CBSR slab [m, k]  --all_gather-->  CBSR of all N nodes
            -> local SpGEMM over the rank's rows (global column ids)          -> out [m, 256]
            5k bytes per node travel instead of 1 KiB (the point of the CBSR format).
            On NCCL the gather is FUSED into the top-k kernel when torch symmetric memory is available
            (gather="peer"): the kernel writes every CBSR row straight into all ranks' gathered buffers over
            NVLink (maxk_topk_cbsr_peers) and one device-side barrier replaces the two all_gather launches.
            Same bytes on the wire; falls back to all_gather (gather="nccl") where peer mapping is unavailable.
backward  : "reduce_scatter" (default): every rank scatters the outer products of ITS rows into a
            full-size partial gs[N, k] (selectors of all nodes are resident since forward), then
            reduce_scatter(sum) -> gs [m, k].      N*k*4 bytes per rank on the wire.
            "allgather" (the variant BASELINE.json names): all_gather the dense gradient rows
            [m, 256] -> [N, 256], then SSpMM over the rank's COLUMN slice A[:, rows_p]; no reduction,
            no cross-rank reduction (the SSpMM itself still accumulates with unordered fp32 atomics, so the
            last bits vary run to run like on one GPU), but 256/k times more bytes on the wire.

The collectives and the partition logic are backend-agnostic (the CPU tests run them on gloo with
world_size 2); the compute backend defaults to the CUDA kernels and there is no CPU fallback in the
product: `compute=` exists so that tests can inject the oracle.
"""
import os

import torch
import torch.distributed as dist

# One NCCL group launch for the two all_gathers of a CBSR slab.  Measured on 8xB200: the coalesced
# launch was SLOWER than two plain all_gathers (Reddit shape, P = 8: 1.60 vs 0.97 ms per step), so it is off.
COALESCE_ALL_GATHER = os.environ.get("MAXK_COALESCE_ALL_GATHER", "0") == "1"


# ------------------------------------------------------------------------------------------------
# partition helpers (pure tensor logic, device agnostic)
# ------------------------------------------------------------------------------------------------
def slab_rows(n, world):
    """Rows per rank (uniform, last slab padded)."""
    return (n + world - 1) // world


def row_bounds(n, world, rank):
    m = slab_rows(n, world)
    lo = min(rank * m, n)
    return lo, min(lo + m, n)


def uniform_bounds(n, world):
    """Slab boundaries [b_0 .. b_P] of the equal-row partition."""
    m = slab_rows(n, world)
    return [min(p * m, n) for p in range(world + 1)]


def balanced_bounds(indptr, world):
    """Slab boundaries with ~E/P edges per rank (prefix sum of the degrees): rank p owns rows
    [b_p, b_p+1).  Equal-row slabs put E/P edges on every rank only on uniform graphs; on a power-law
    graph the rank that holds the hubs sets the step time."""
    indptr = torch.as_tensor(indptr).long().cpu()
    n, e = indptr.numel() - 1, int(indptr[-1])
    targets = torch.tensor([(p * e) // world for p in range(1, world)], dtype=torch.int64)
    cuts = torch.searchsorted(indptr, targets, right=False).clamp(max=n).tolist() if world > 1 else []
    bounds = [0] + cuts + [n]
    for p in range(1, len(bounds)):                   # monotone (empty slabs are legal)
        bounds[p] = max(bounds[p], bounds[p - 1])
    return bounds


def padded_position(ids, bounds, m):
    """Position of global row / column ids in the padded numbering owner * m + (id - b_owner), the layout
    of every all_gathered [P * m, ...] buffer.  Identity for the equal-row partition."""
    b = torch.as_tensor(bounds[:-1], dtype=torch.int64, device=ids.device)
    ids = ids.long()
    owner = torch.searchsorted(b, ids, right=True) - 1
    # rows of empty slabs share a boundary: searchsorted(right=True) picks the last slab starting there,
    # which is the one that owns the row
    return owner * m + (ids - b[owner])


def shard_rows(graph, world, rank, bounds=None):
    """Row slab of a CSR graph: local indptr (rebased, padded to m rows), column ids in the padded global
    numbering (== global ids for the equal-row partition)."""
    n = graph["v_num"]
    uniform = bounds is None
    if uniform:
        bounds = uniform_bounds(n, world)
    m = max(1, max(bounds[p + 1] - bounds[p] for p in range(world)))
    lo, hi = bounds[rank], bounds[rank + 1]
    indptr = graph["indptr"]
    e0, e1 = int(indptr[lo]), int(indptr[hi])
    local_ptr = torch.full((m + 1,), e1 - e0, dtype=torch.int32, device=indptr.device)
    local_ptr[: hi - lo + 1] = indptr[lo:hi + 1] - e0
    indices = graph["indices"][e0:e1]
    if not uniform:
        indices = padded_position(indices, bounds, m).to(torch.int32)
    return {"indptr": local_ptr, "indices": indices.contiguous(),
            "values": graph["values"][e0:e1].contiguous(), "v_num": m, "e_num": e1 - e0,
            "n_global": n, "row_lo": lo, "row_hi": hi}


def shard_columns(graph, world, rank, bounds=None):
    """Column slab A[:, rows_p] as a CSR over ALL source rows in the padded numbering (P*m rows), local
    column ids.  Used by the all_gather backward variant."""
    n = graph["v_num"]
    if bounds is None:
        bounds = uniform_bounds(n, world)
    m = max(1, max(bounds[p + 1] - bounds[p] for p in range(world)))
    lo, hi = bounds[rank], bounds[rank + 1]
    indptr, indices, values = graph["indptr"].long(), graph["indices"], graph["values"]
    keep = (indices >= lo) & (indices < hi)
    rows = torch.repeat_interleave(torch.arange(n, device=indices.device), indptr[1:] - indptr[:-1])
    counts = torch.bincount(padded_position(rows[keep], bounds, m), minlength=world * m)
    ptr = torch.zeros(world * m + 1, dtype=torch.int64, device=indices.device)
    ptr[1:] = torch.cumsum(counts, 0)
    return {"indptr": ptr.to(torch.int32), "indices": (indices[keep] - lo).to(torch.int32).contiguous(),
            "values": values[keep].contiguous(), "v_num": world * m, "e_num": int(keep.sum())}


# ------------------------------------------------------------------------------------------------
# collectives with uniform counts
# ------------------------------------------------------------------------------------------------
def _all_gather(local, group):
    world = dist.get_world_size(group)
    out = local.new_empty((world * local.size(0),) + tuple(local.shape[1:]))
    dist.all_gather_into_tensor(out, local.contiguous(), group=group)
    return out


def _all_gather_pair(a, b, group):
    """Two all_gathers issued as ONE NCCL group launch (values + selectors of the CBSR slab)."""
    world = dist.get_world_size(group)
    out_a = a.new_empty((world * a.size(0),) + tuple(a.shape[1:]))
    out_b = b.new_empty((world * b.size(0),) + tuple(b.shape[1:]))
    if COALESCE_ALL_GATHER and dist.get_backend(group) == "nccl" and hasattr(dist, "_coalescing_manager"):
        with dist._coalescing_manager(group=group, device=a.device, async_ops=False):
            dist.all_gather_into_tensor(out_a, a.contiguous(), group=group)
            dist.all_gather_into_tensor(out_b, b.contiguous(), group=group)
    else:
        dist.all_gather_into_tensor(out_a, a.contiguous(), group=group)
        dist.all_gather_into_tensor(out_b, b.contiguous(), group=group)
    return out_a, out_b


def _reduce_scatter_sum(full, group):
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    m = full.size(0) // world
    out = full.new_empty((m,) + tuple(full.shape[1:]))
    if dist.get_backend(group) == "gloo":       # gloo has no reduce_scatter; only the CPU tests come here
        dist.all_reduce(full, group=group)
        out.copy_(full[rank * m:(rank + 1) * m])
    else:
        dist.reduce_scatter_tensor(out, full.contiguous(), op=dist.ReduceOp.SUM, group=group)
    return out


# ------------------------------------------------------------------------------------------------
# compute backend: the CUDA kernels
# ------------------------------------------------------------------------------------------------
class CudaCompute:
    """The product backend: libmaxk_b200.so through maxk_cuda_kernels (raises if it is missing)."""

    def __init__(self):
        import maxk_cuda_kernels
        self.k = maxk_cuda_kernels

    def topk(self, x, k):
        r = self.k.topk_cbsr(x, k, order=self.k.ORDER_BANKED)
        return r["values"], r["sel"]

    def topk_masked(self, x, k):
        """(values, selectors, x with every non-selected entry zeroed) in one pass."""
        r = self.k.topk_cbsr(x, k, order=self.k.ORDER_BANKED, want_masked=True)
        return r["values"], r["sel"], r["masked"]

    def mask_apply(self, dense, sel, add_vals):
        """dense * mask(sel) + scatter(add_vals): the MaxK backward plus the aggregation gradient, one kernel."""
        return self.k.mask_apply(dense, sel, add_vals=add_vals)

    def spgemm(self, g, vals, sel, row_div=None):
        ip = g["indptr"]
        if g.get("plan") is None:                 # the slab's row plan, built once
            g["plan"] = self.k.build_plan(ip[:-1], ip[1:])
        return self.k.spgemm_forward_csr(ip[:-1], ip[1:], g["indices"], g["values"], vals, sel, row_div=row_div,
                                         plan=g["plan"])

    def sspmm(self, g, grad, sel, row_div=None):
        ip = g["indptr"]
        return self.k.sspmm_backward_csr(ip[:-1], ip[1:], g["indices"], g["values"], grad, sel, row_div=row_div)


def _multicast_ptr(handle):
    """NVLS multicast address of a symmetric-memory allocation, or 0 when the system has none."""
    try:
        if not handle.has_multicast_support:
            return 0
        return int(handle.multicast_ptr or 0)
    except Exception:
        return 0


class PeerReduce:
    """The backward's partial sampled gradient [P*m, k] in symmetric memory, two sets (call c uses set c % 2): after
    one barrier every rank sums ITS slab over all ranks with multimem.ld_reduce -- the reduction happens inside the
    NVSwitch, a rank receives m*k*4 bytes instead of reading (P-1) slabs.  The barrier of call c+1 orders every
    rank's read of set c % 2 before anybody zero-fills it again in call c+2."""

    def __init__(self, rows_total, k, device, group):
        import torch.distributed._symmetric_memory as symm
        pg = group if group is not None else dist.group.WORLD
        self.sets = []
        for _ in range(2):
            buf = symm.empty((rows_total, k), dtype=torch.float32, device=device)
            h = symm.rendezvous(buf, pg)
            mc = _multicast_ptr(h)
            if not mc:
                raise RuntimeError("no NVLS multicast mapping for symmetric memory on this system")
            self.sets.append({"buf": buf, "h": h, "mc": mc})
        self.call = 0

    def next_set(self):
        s = self.sets[self.call % 2]
        self.call += 1
        return s


class PeerGather:
    """Gathered CBSR buffers [P*m, k] in torch symmetric memory (peer-mapped over NVLink), two sets: step s uses
    set s % 2.  One barrier per step (after the peer writes) is enough: a rank can only be one barrier ahead of
    the slowest one, so nobody writes set s % 2 again before every rank has finished reading it."""

    def __init__(self, rows_total, k, device, group, use_multicast=True):
        import torch.distributed._symmetric_memory as symm
        pg = group if group is not None else dist.group.WORLD
        self.sets = []
        for _ in range(2):
            vals = symm.empty((rows_total, k), dtype=torch.float32, device=device)
            sel = symm.empty((rows_total, k), dtype=torch.uint8, device=device)
            hv, hs = symm.rendezvous(vals, pg), symm.rendezvous(sel, pg)
            mc = use_multicast and _multicast_ptr(hv) and _multicast_ptr(hs)
            self.sets.append({"vals": vals, "sel": sel, "hv": hv, "hs": hs,
                              "val_ptrs": list(hv.buffer_ptrs), "sel_ptrs": list(hs.buffer_ptrs),
                              "mc_val": _multicast_ptr(hv) if mc else 0, "mc_sel": _multicast_ptr(hs) if mc else 0})
        self.multicast = bool(self.sets[0]["mc_val"])
        self.step = 0

    def next_set(self):
        s = self.sets[self.step % 2]
        self.step += 1
        return s


class ShardedMaxKAggregation:
    """top-k -> all_gather(CBSR) -> SpGEMM, and its backward, for one rank's row slab."""

    def __init__(self, graph, k, group=None, backward_mode="reduce_scatter", compute=None, row_div=None,
                 partition="rows", gather="auto"):
        if backward_mode not in ("reduce_scatter", "allgather"):
            raise ValueError("backward_mode must be 'reduce_scatter' or 'allgather'")
        if gather not in ("auto", "peer", "nccl"):
            raise ValueError("gather must be 'auto', 'peer' (top-k writes into peer memory) or 'nccl' (all_gather)")
        if partition not in ("rows", "nnz"):
            raise ValueError("partition must be 'rows' (equal row slabs) or 'nnz' (equal edge counts)")
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.k = int(k)
        self.n = graph["v_num"]
        self.partition = partition
        # every rank derives the same boundaries from the same indptr: nothing is exchanged
        self.bounds = uniform_bounds(self.n, self.world) if partition == "rows" else balanced_bounds(graph["indptr"], self.world)
        explicit = None if partition == "rows" else self.bounds
        self.rows = shard_rows(graph, self.world, self.rank, explicit)
        self.m = self.rows["v_num"]                   # padded slab height, the same on every rank
        self.cols = shard_columns(graph, self.world, self.rank, explicit) if backward_mode == "allgather" else None
        self.backward_mode = backward_mode
        self.compute = compute if compute is not None else CudaCompute()
        self.row_div = None
        if row_div is not None:                       # per-row divisor of the full graph -> local slab (pad with 1)
            lo, hi = self.rows["row_lo"], self.rows["row_hi"]
            self.row_div = torch.ones(self.m, dtype=torch.float32, device=row_div.device)
            self.row_div[: hi - lo] = row_div[lo:hi]
        self.sel_full = None
        # forward exchange: fused into the top-k kernel over peer-mapped memory when possible
        self.peer = None
        self.gather_error = None
        cuda_nccl = isinstance(self.compute, CudaCompute) and self.rows["indices"].is_cuda and \
            dist.get_backend(group) == "nccl" and self.world <= self.compute.k.MAX_PEERS and \
            self.compute.k._lib.maxk_banked_modulus(self.k) >= 4
        multicast = os.environ.get("MAXK_NVLS", "1") == "1"
        if gather != "nccl" and cuda_nccl and os.environ.get("MAXK_PEER_GATHER", "1") == "1":
            try:
                self.peer = PeerGather(self.world * self.m, self.k, self.rows["indices"].device, group, multicast)
            except Exception as ex:          # no peer mapping on this system: NCCL all_gather does the same job
                self.gather_error = repr(ex)[:300]
            # every rank must take the same path (a rank alone at a barrier would hang the others)
            if not self._all_ranks(self.peer is not None):
                self.peer = None
                if gather == "peer":
                    raise RuntimeError("gather='peer' needs peer-mapped symmetric memory on every rank: %s" % self.gather_error)
            elif not self._all_ranks(self.peer.multicast):
                for st in self.peer.sets:
                    st["mc_val"] = st["mc_sel"] = 0
                self.peer.multicast = False
        self.gather = ("peer+nvls" if self.peer.multicast else "peer") if self.peer is not None else "nccl"
        # backward exchange: reduce_scatter inside the NVSwitch when the partial can live in multicast-mapped memory
        # (small partials only: the switch pulls every rank's WHOLE partial, own slab included, so above a few tens of
        # MB it is link-bound at ~P/(P-1) times the bytes of a reduce_scatter -- products shape, 4 GPUs: 0.47 ms for a
        # 313 MB partial -- while below it wins on latency: Reddit shape 48 us against 68 us)
        self.peer_reduce = None
        small = self.world * self.m * self.k * 4 <= int(os.environ.get("MAXK_NVLS_REDUCE_MAX_MB", "64")) << 20
        if self.peer is not None and multicast and small and backward_mode == "reduce_scatter" and (self.m * self.k) % 4 == 0:
            try:
                self.peer_reduce = PeerReduce(self.world * self.m, self.k, self.rows["indices"].device, group)
            except Exception as ex:
                self.gather_error = (self.gather_error or "") + " | backward: " + repr(ex)[:200]
            if not self._all_ranks(self.peer_reduce is not None):
                self.peer_reduce = None
        self.reduce = "nvls" if self.peer_reduce is not None else ("nccl" if backward_mode == "reduce_scatter" else "none")

    def _all_ranks(self, ok):
        """True iff `ok` holds on every rank of the group."""
        dev = self.rows["indices"].device
        flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.group)
        return bool(flag.item())

    def local_slab(self, full):
        """This rank's rows of a full [N, ...] tensor, zero-padded to the slab height m."""
        lo, hi = self.rows["row_lo"], self.rows["row_hi"]
        out = full.new_zeros((self.m,) + tuple(full.shape[1:]))
        out[: hi - lo] = full[lo:hi]
        return out

    def valid_rows(self):
        """Number of real (non-padding) rows of this rank's slab."""
        return self.rows["row_hi"] - self.rows["row_lo"]

    def gather_cbsr(self, x_local, want_masked=False):
        """top-k of this rank's slab + exchange -> (vals_full [P*m, k], sel_full [P*m, k], masked slab or None)."""
        if self.peer is not None and x_local.size(1) == 256:
            st = self.peer.next_set()
            masked = self.compute.k.topk_cbsr_to_peers(x_local, self.k, st["val_ptrs"], st["sel_ptrs"],
                                                       row_offset=self.rank * self.m, want_masked=want_masked,
                                                       mc_val_ptr=st["mc_val"], mc_sel_ptr=st["mc_sel"])
            st["hv"].barrier(channel=0)          # every rank's rows have landed in every rank's buffers
            return st["vals"], st["sel"], masked
        masked = None
        if want_masked and hasattr(self.compute, "topk_masked"):
            vals, sel, masked = self.compute.topk_masked(x_local, self.k)
        else:
            vals, sel = self.compute.topk(x_local, self.k)
            if want_masked:                      # injected test backends: plain torch
                masked = torch.zeros_like(x_local).scatter_(1, sel.long(), vals)
        vals_full, sel_full = _all_gather_pair(vals, sel, self.group)
        return vals_full, sel_full, masked

    # x_local: [m, 256] (rows past the end of the slab are padding and may hold anything finite)
    def forward(self, x_local):
        vals_full, self.sel_full, _ = self.gather_cbsr(x_local)
        return self.compute.spgemm(self.rows, vals_full, self.sel_full, self.row_div)

    def backward(self, grad_local):
        if self.sel_full is None:
            raise RuntimeError("backward() before forward()")
        if self.backward_mode == "reduce_scatter":
            if self.peer_reduce is not None and grad_local.is_cuda:
                st = self.peer_reduce.next_set()
                ip = self.rows["indptr"]
                self.compute.k.sspmm_backward_csr(ip[:-1], ip[1:], self.rows["indices"], self.rows["values"], grad_local,
                                                  self.sel_full, row_div=self.row_div, out=st["buf"])
                st["h"].barrier(channel=0)               # every rank's partial is complete
                gs = torch.empty(self.m, self.k, dtype=torch.float32, device=grad_local.device)
                return self.compute.k.nvls_reduce(st["mc"] + self.rank * self.m * self.k * 4, gs)
            partial = self.compute.sspmm(self.rows, grad_local, self.sel_full, self.row_div)   # [P*m, k]
            return _reduce_scatter_sum(partial, self.group)
        grad = grad_local if self.row_div is None else grad_local / self.row_div.unsqueeze(-1)
        grad_full = _all_gather(grad, self.group)                                             # [P*m, 256]
        lo = self.rank * self.m
        return self.compute.sspmm(self.cols, grad_full, self.sel_full[lo:lo + self.m].contiguous())

    def wire_bytes(self):
        """Bytes each rank RECEIVES per forward / backward (for the report)."""
        others = (self.world - 1) * self.m
        fwd = others * self.k * 5
        bwd = others * 256 * 4 if self.backward_mode == "allgather" else others * self.k * 4
        return {"forward": fwd, "backward": bwd}


class _ShardedFn(torch.autograd.Function):
    """x_local -> (x_local * topk_mask, A[rows_p, :] @ maxk(x)) with the selectors kept in ctx, so one
    ShardedMaxKAggregation can serve every layer of a model."""

    @staticmethod
    def forward(ctx, x_local, layer):
        vals_full, sel_full, masked = layer.gather_cbsr(x_local, want_masked=True)
        out = layer.compute.spgemm(layer.rows, vals_full, sel_full, layer.row_div)
        if layer.peer is not None:
            sel_full = sel_full.clone()          # the peer buffers are recycled two steps later; autograd keeps its own
        ctx.layer = layer
        ctx.save_for_backward(sel_full)
        layer.sel_full = sel_full
        return masked, out

    @staticmethod
    def backward(ctx, grad_masked, grad_out):
        layer = ctx.layer
        (sel_full,) = ctx.saved_tensors
        lo = layer.rank * layer.m
        sel_local = sel_full[lo:lo + layer.m].contiguous()
        layer.sel_full = sel_full
        gs = layer.backward(grad_out.contiguous())
        if hasattr(layer.compute, "mask_apply"):              # grad * mask + scatter(gs) in one kernel
            return layer.compute.mask_apply(grad_masked.contiguous(), sel_local, gs), None
        dense = torch.zeros(gs.size(0), grad_masked.size(1), dtype=gs.dtype, device=gs.device)
        dense.scatter_(1, sel_local.long(), gs)
        mask = torch.zeros_like(grad_masked).scatter_(1, sel_local.long(), 1.0)
        return grad_masked * mask + dense, None


def sharded_maxk_act_spgemm(x_local, layer):
    """Autograd entry point: (maxk(x_local), A[rows_p, :] @ maxk(x)) with gradients flowing to x_local."""
    return _ShardedFn.apply(x_local, layer)


def sharded_maxk_spgemm(x_local, layer):
    """Aggregation only (the masked features are dropped)."""
    return _ShardedFn.apply(x_local, layer)[1]


# ------------------------------------------------------------------------------------------------
# Row-sharded GraphSAGE (BASELINE.json config 5: "GraphSAGE MaxK k=32 row-sharded at 2/4/8 B200 with
# NCCL allgather of CBSR rows").  Same formulas as MaxKSAGE / MaxKSAGEConv (model_integrated_v3.py:62-192,
# 522-591); every rank holds its row slab of the features and a full replica of the weights, whose
# gradients are summed with one all_reduce per step (allreduce_gradients).
# ------------------------------------------------------------------------------------------------
class ShardedMaxKSAGE(torch.nn.Module):
    def __init__(self, graph, in_size, hid_size, num_hid_layers, out_size, maxk=32, feat_drop=0.0, norm=False,
                 group=None, compute=None, backward_mode="reduce_scatter", partition="rows"):
        super().__init__()
        nn = torch.nn
        deg = torch.clamp((graph["indptr"][1:] - graph["indptr"][:-1]).to(torch.float32), min=1.0)
        self.agg = ShardedMaxKAggregation(graph, maxk, group=group, backward_mode=backward_mode, compute=compute,
                                          row_div=deg, partition=partition)
        self.group = group
        self.fc_self = nn.ModuleList(nn.Linear(hid_size, hid_size) for _ in range(num_hid_layers))
        self.fc_neigh = nn.ModuleList(nn.Linear(hid_size, hid_size, bias=False) for _ in range(num_hid_layers))
        self.norms = nn.ModuleList(nn.LayerNorm(hid_size) for _ in range(num_hid_layers if norm else 0))
        self.drop = nn.Dropout(feat_drop)
        self.lin_in = nn.Linear(in_size, hid_size)
        self.lin_out = nn.Linear(hid_size, out_size)
        gain = nn.init.calculate_gain("relu")
        for a, b in zip(self.fc_self, self.fc_neigh):
            nn.init.xavier_uniform_(a.weight, gain=gain)
            nn.init.xavier_uniform_(b.weight, gain=gain)
        nn.init.xavier_uniform_(self.lin_in.weight)
        nn.init.xavier_uniform_(self.lin_out.weight)

    def forward(self, x_local):
        """x_local: this rank's row slab [m, in_size] (rows past the end of the graph are padding)."""
        h = self.lin_in(x_local)
        for i in range(len(self.fc_self)):
            h_sparse, h_agg = sharded_maxk_act_spgemm(h, self.agg)
            h = self.fc_self[i](self.drop(h_sparse)) + self.fc_neigh[i](h_agg)
            if len(self.norms):
                h = self.norms[i](h)
        return self.lin_out(h)


def allreduce_gradients(module, group=None):
    """Sum the weight gradients over the ranks (each rank back-propagated the loss of its own rows)."""
    grads = [p.grad for p in module.parameters() if p.grad is not None]
    if not grads:
        return
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, group=group)
    off = 0
    for g in grads:
        g.copy_(flat[off:off + g.numel()].view_as(g))
        off += g.numel()
