"""MaxK nonlinearity -- the autograd functions the reference's models call before aggregation.

Same names and call shape as the reference:
    MaxK.apply(x, k)      -> x * topk_mask                     (maxk_models_integrated.py:28-43,
                                                                 utils/models.py:11-25)
    OPTMaxK.apply(x, k)   -> (x * topk_mask, topk_values, topk_indices)
                                                                (model_integrated_v3.py:28-43)
but one fused CUDA pass (exact top-k + masked row + CBSR emission) instead of
topk + zeros_like + scatter_ + multiply, and the state kept for backward is the uint8 selector
[N, k] instead of an fp32 [N, 256] mask.

The conv layers / models of the reference (model_integrated_v3.py:62-752) follow below, on raw CSR tensors
instead of a DGL graph.
"""
import torch
from torch.autograd import Function

import maxk_cuda_kernels as _k


def _require_cuda_2d(x, who):
    if not (isinstance(x, torch.Tensor) and x.is_cuda and x.dim() == 2):
        raise RuntimeError("%s: input must be a 2-D CUDA tensor (there is no CPU fallback)" % who)
    if x.dtype != torch.float32:
        raise RuntimeError("%s: input must be float32" % who)
    if x.size(1) > _k.FULL_DIM:
        raise RuntimeError("%s: feature dim %d > 256 cannot be addressed by uint8 selectors" % (who, x.size(1)))


class MaxK(Function):
    """Standard MaxK activation (maxk_models_integrated.py:28-43)."""

    @staticmethod
    def forward(ctx, input, k=1):
        _require_cuda_2d(input, "MaxK")
        r = _k.topk_cbsr(input, int(k), order=_k.ORDER_BANKED, want_masked=True)
        ctx.save_for_backward(r["sel"])
        return r["masked"]

    @staticmethod
    def backward(ctx, grad_output):
        (sel,) = ctx.saved_tensors
        return _k.mask_apply(grad_output.contiguous(), sel), None


class OPTMaxK(Function):
    """MaxK activation that also returns the CBSR pair (model_integrated_v3.py:28-43).

    topk_indices is int64 like torch.topk's (it may be fed to scatter_/gather by callers).  Entry order inside a
    row: `OPTMaxK.order = "banked"` (default, the order the SpGEMM kernel is fastest on; every consumer on this
    path -- spmm, scatter -- is order-independent) or `"value_desc"`, torch.topk's own order (values descending,
    lowest column first on ties), i.e. exactly what the reference class returns.

    In-tree callers pass uint8_indices=True and get the kernel's uint8 selectors instead of an int64 copy (the
    aggregation operator needs uint8; the int64 round trip costs N*k*9 bytes of traffic per layer).

    reference_compat: the reference's backward returns grad_output * mask only and DROPS
    grad_topk_values (model_integrated_v3.py:40-43), so the aggregation branch
    never back-propagates into x.  Default here is the correct gradient
    (grad_output * mask + scatter(grad_topk_values)); set OPTMaxK.reference_compat = True to
    reproduce the reference bit for bit.
    """
    reference_compat = False
    order = "banked"

    @staticmethod
    def forward(ctx, input, k=1, uint8_indices=False):
        _require_cuda_2d(input, "OPTMaxK")
        if OPTMaxK.order not in ("banked", "value_desc"):
            raise ValueError('OPTMaxK.order must be "banked" or "value_desc"')
        order = _k.ORDER_BANKED if OPTMaxK.order == "banked" else _k.ORDER_VALUE_DESC
        r = _k.topk_cbsr(input, int(k), order=order, want_masked=True, want_i64=not uint8_indices)
        ctx.save_for_backward(r["sel"])
        idx = r["sel"] if uint8_indices else r["i64"]
        ctx.mark_non_differentiable(idx)
        return r["masked"], r["values"], idx

    @staticmethod
    def backward(ctx, grad_output, grad_topk_values, grad_topk_indices):
        (sel,) = ctx.saved_tensors
        add = None
        if grad_topk_values is not None and not OPTMaxK.reference_compat:
            add = grad_topk_values.contiguous()
        return _k.mask_apply(grad_output.contiguous(), sel, add), None, None



# ==================================================================================================
# Conv layers and models: the callers either side of the hot path, without DGL.
#
# Same class names, constructor arguments and forward formulas as the reference generation that
# really reaches the kernels (model_integrated_v3.py: MaxKSAGEConv :62-192, MaxKGraphConv :194-398,
# MaxKGINConv :400-520, MaxKSAGE/MaxKGCN/MaxKGIN :522-752), but the `graph` argument is a CSRGraph
# (raw CSR tensors) instead of a DGL graph, and the warp4 metadata is built on the GPU instead of
# being read from kernels/w12_nz64_warp_4/<graph>.warp4.  Dense parts (Linear, LayerNorm, dropout)
# are plain PyTorch/cuBLAS like in the reference.
#
# Reference semantics kept on purpose: the aggregation operator always divides by
# the clamped degree, also for GCN norm="both" and for GIN "sum"; GCN's left normalisation acts on
# the dense masked features, which the MaxK branch does not use.
# ==================================================================================================
import torch.nn as nn
import torch.nn.init as init

from spgemmfunction_v4 import MaxKSpmmWrapper


class CSRGraph:
    """Undirected graph as CSR tensors on one device (what graph.adj_tensors('csr') gives the reference,
    model_integrated_v3.py:113)."""

    def __init__(self, indptr, indices, graph_name="graph"):
        self.indptr = indptr.to(torch.int32).contiguous()
        self.indices = indices.to(torch.int32).contiguous()
        self.graph_name = graph_name
        self.device = indices.device
        self.values = torch.ones(self.indices.numel(), device=self.device, dtype=torch.float32)   # :129
        deg = (self.indptr[1:] - self.indptr[:-1]).to(torch.float32)
        self.degrees = torch.clamp(deg, min=1.0)                                                    # :116-117
        self._wrapper = None

    @classmethod
    def from_dict(cls, g, graph_name="graph"):
        return cls(g["indptr"], g["indices"], graph_name)

    def num_nodes(self):
        return self.indptr.numel() - 1

    def num_edges(self):
        return self.indices.numel()

    def in_degrees(self):
        return self.indptr[1:] - self.indptr[:-1]

    out_degrees = in_degrees          # undirected graphs only, as the reference assumes (spgemmfunction_v4:4)

    def wrapper(self):
        """One MaxKSpmmWrapper (warp4 metadata) per graph, shared by all layers."""
        if self._wrapper is None:
            self._wrapper = MaxKSpmmWrapper(self.graph_name)
            self._wrapper.build_metadata(self.indptr)
        return self._wrapper

    def aggregate(self, topk_values, topk_indices):
        """(A @ scatter(topk)) / degrees through the MaxK kernels (autograd-aware)."""
        return self.wrapper().spmm(self.indices, self.values, topk_values, topk_indices, self.indptr, self.degrees)


class MaxKSAGEConv(nn.Module):
    """model_integrated_v3.py:62-192 (mean aggregator): rst = fc_self(h) + fc_neigh(mean_agg(topk(h)))."""

    def __init__(self, in_feats, out_feats, aggregator_type="mean", feat_drop=0., bias=True, norm=None,
                 activation=None, k_value=32):
        super().__init__()
        if aggregator_type != "mean":
            raise ValueError("Only 'mean' supported, got %s" % aggregator_type)
        self._in_src_feats = self._in_dst_feats = in_feats
        self._out_feats = out_feats
        self.norm = norm
        self.feat_drop = nn.Dropout(feat_drop)
        self.activation = activation
        self.k_value = k_value
        self.fc_neigh = nn.Linear(in_feats, out_feats, bias=False)
        self.fc_self = nn.Linear(in_feats, out_feats, bias=bias)
        self.reset_parameters()

    def reset_parameters(self):
        gain = nn.init.calculate_gain("relu")
        nn.init.xavier_uniform_(self.fc_self.weight, gain=gain)
        nn.init.xavier_uniform_(self.fc_neigh.weight, gain=gain)

    def forward(self, graph, feat, topk_values=None, topk_indices=None):
        if topk_values is None or topk_indices is None:
            raise RuntimeError("topk_values / topk_indices REQUIRED")               # :149-150
        # in_feats > out_feats: the reference switches to transform-before-aggregate here (:161-172), but applies
        # fc_neigh to the [N, k] top-k VALUES, which only type-checks when k == in_feats == out_feats.  The order is
        # a FLOP optimisation, not a different function -- A (X_s W) == (A X_s) W -- and only the k-sparse X_s can go
        # through the CBSR kernels, so both cases aggregate first and transform afterwards.
        h_self = self.feat_drop(feat)
        h_neigh = self.fc_neigh(graph.aggregate(topk_values, topk_indices))           # :175-182
        rst = self.fc_self(h_self) + h_neigh                                          # :185
        if self.activation is not None:
            rst = self.activation(rst)
        if self.norm is not None:
            rst = self.norm(rst)
        return rst


class MaxKGraphConv(nn.Module):
    """model_integrated_v3.py:194-398."""

    def __init__(self, in_feats, out_feats, norm="both", weight=True, bias=True, activation=None,
                 allow_zero_in_degree=False, k_value=32):
        super().__init__()
        if norm not in ("none", "both", "right", "left"):
            raise ValueError('Invalid norm value. Must be either "none", "both", "right" or "left". But got "%s".' % norm)
        self._in_feats, self._out_feats, self._norm = in_feats, out_feats, norm
        self._allow_zero_in_degree = allow_zero_in_degree
        self.k_value = k_value
        self.weight = nn.Parameter(torch.empty(in_feats, out_feats)) if weight else None
        self.bias = nn.Parameter(torch.empty(out_feats)) if bias else None
        self._activation = activation
        self.reset_parameters()

    def reset_parameters(self):
        if self.weight is not None:
            init.xavier_uniform_(self.weight)
        if self.bias is not None:
            init.zeros_(self.bias)

    def forward(self, graph, feat, topk_values=None, topk_indices=None):
        if topk_values is None or topk_indices is None:
            raise RuntimeError("topk_values / topk_indices REQUIRED (the DGL fallback of the reference is not part of this package)")
        if not self._allow_zero_in_degree and bool((graph.in_degrees() == 0).any()):
            raise ValueError("There are 0-in-degree nodes in the graph, output for those nodes will be invalid.")  # :281-290
        rst = graph.aggregate(topk_values, topk_indices)                              # :341-346 (aggregate then transform)
        if self.weight is not None:
            rst = torch.matmul(rst, self.weight)                                       # :327-349; A (X_s W) == (A X_s) W, see MaxKSAGEConv
        if self._norm in ("right", "both"):                                           # :378-386
            degs = graph.in_degrees().to(rst).clamp(min=1)
            rst = rst * (torch.pow(degs, -0.5) if self._norm == "both" else 1.0 / degs).unsqueeze(-1)
        if self.bias is not None:
            rst = rst + self.bias
        if self._activation is not None:
            rst = self._activation(rst)
        return rst


class MaxKGINConv(nn.Module):
    """model_integrated_v3.py:400-520 (sum aggregator): rst = (1 + eps) * h + agg(topk(h))."""

    def __init__(self, apply_func=None, aggregator_type="sum", init_eps=0, learn_eps=False, activation=None, k_value=32):
        super().__init__()
        if aggregator_type != "sum":
            raise KeyError("Only the 'sum' aggregator reaches the MaxK kernels (model_integrated_v3.py:457-458)")
        self.apply_func, self.activation, self.k_value = apply_func, activation, k_value
        if learn_eps:
            self.eps = nn.Parameter(torch.FloatTensor([init_eps]))
        else:
            self.register_buffer("eps", torch.FloatTensor([init_eps]))

    def forward(self, graph, feat, topk_values=None, topk_indices=None):
        if topk_values is None or topk_indices is None:
            raise RuntimeError("topk_values / topk_indices REQUIRED")
        rst = (1 + self.eps) * feat + graph.aggregate(topk_values, topk_indices)     # :487-507
        if self.apply_func is not None:
            rst = self.apply_func(rst)
        if self.activation is not None:
            rst = self.activation(rst)
        return rst


class MaxKSAGE(nn.Module):
    """model_integrated_v3.py:522-591."""

    def __init__(self, in_size, hid_size, num_hid_layers, out_size, maxk=32, feat_drop=0.5, norm=False,
                 nonlinear="maxk", graph_name=""):
        super().__init__()
        if nonlinear != "maxk":
            raise ValueError("Only 'maxk' supported, got %s" % nonlinear)
        self.num_layers, self.graph_name, self.k_value = num_hid_layers, graph_name, maxk
        self.layers = nn.ModuleList(
            MaxKSAGEConv(hid_size, hid_size, "mean", feat_drop,
                         norm=nn.LayerNorm(hid_size, elementwise_affine=True) if norm else None, k_value=maxk)
            for _ in range(num_hid_layers))
        self.lin_in = nn.Linear(in_size, hid_size)
        self.lin_out = nn.Linear(hid_size, out_size)
        init.xavier_uniform_(self.lin_in.weight)
        init.xavier_uniform_(self.lin_out.weight)

    def forward(self, g, x):
        x = self.lin_in(x)
        for layer in self.layers:
            x_sparse, topk_values, topk_indices = OPTMaxK.apply(x, self.k_value, True)      # :581
            x = layer(g, x_sparse, topk_values, topk_indices)
        return self.lin_out(x)


class _MaxKStack(nn.Module):
    """Shared body of MaxKGCN (:593-672) and MaxKGIN (:674-752): Linear -> MaxK -> dropout -> conv [-> LayerNorm]."""

    def __init__(self, conv_factory, in_size, hid_size, num_hid_layers, out_size, maxk, feat_drop, norm, nonlinear,
                 graph_name):
        super().__init__()
        if nonlinear != "maxk":
            raise ValueError("only nonlinear='maxk' reaches the MaxK kernels")
        self.num_layers, self.norm_flag, self.graph_name, self.k_value = num_hid_layers, norm, graph_name, maxk
        self.dropoutlayers = nn.ModuleList(nn.Dropout(feat_drop) for _ in range(num_hid_layers))
        self.convlayers = nn.ModuleList(conv_factory() for _ in range(num_hid_layers))
        self.normlayers = nn.ModuleList(nn.LayerNorm(hid_size, elementwise_affine=True)
                                        for _ in range(num_hid_layers if norm else 0))
        self.linlayers = nn.ModuleList(nn.Linear(hid_size, hid_size) for _ in range(num_hid_layers))
        for lin in self.linlayers:
            init.xavier_uniform_(lin.weight)
        self.lin_in = nn.Linear(in_size, hid_size)
        self.lin_out = nn.Linear(hid_size, out_size)
        init.xavier_uniform_(self.lin_in.weight)
        init.xavier_uniform_(self.lin_out.weight)

    def forward(self, g, x):
        x = self.lin_in(x).relu()
        for i in range(self.num_layers):
            x = self.linlayers[i](x)
            x_sparse, topk_values, topk_indices = OPTMaxK.apply(x, self.k_value, True)
            x_sparse = self.dropoutlayers[i](x_sparse)
            x = self.convlayers[i](g, x_sparse, topk_values, topk_indices)
            if self.norm_flag:
                x = self.normlayers[i](x)
        return self.lin_out(x)


class MaxKGCN(_MaxKStack):
    def __init__(self, in_size, hid_size, num_hid_layers, out_size, maxk=32, feat_drop=0.5, norm=False,
                 nonlinear="maxk", graph_name=""):
        super().__init__(lambda: MaxKGraphConv(hid_size, hid_size, norm="both", weight=False, bias=False, k_value=maxk),
                         in_size, hid_size, num_hid_layers, out_size, maxk, feat_drop, norm, nonlinear, graph_name)


class MaxKGIN(_MaxKStack):
    def __init__(self, in_size, hid_size, num_hid_layers, out_size, maxk=32, feat_drop=0.5, norm=False,
                 nonlinear="maxk", graph_name=""):
        super().__init__(lambda: MaxKGINConv(None, "sum", init_eps=0, learn_eps=True, k_value=maxk),
                         in_size, hid_size, num_hid_layers, out_size, maxk, feat_drop, norm, nonlinear, graph_name)
