"""MaxK nonlinearity -- the autograd functions the reference's models call before aggregation.

Same names and call shape as the reference:
    MaxK.apply(x, k)      -> x * topk_mask                     (maxk_models_integrated.py:28-43,
                                                                 utils/models.py:11-25)
    OPTMaxK.apply(x, k)   -> (x * topk_mask, topk_values, topk_indices)
                                                                (model_integrated_v3.py:28-43)
but one fused CUDA pass (exact top-k + masked row + CBSR emission) instead of
topk + zeros_like + scatter_ + multiply, and the state kept for backward is the uint8 selector
[N, k] instead of an fp32 [N, 256] mask.

The conv layers / models of the reference file are callers of this path and
are not part of this package.
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

    topk_indices is int64 like torch.topk's (it may be fed to scatter_/gather by callers); the
    entries of a row are in bank-residue-major column order, not value order -- every consumer on this path
    (spmm, scatter) is order-independent.

    reference_compat: the reference's backward returns grad_output * mask only and DROPS
    grad_topk_values (model_integrated_v3.py:40-43), so the aggregation branch
    never back-propagates into x.  Default here is the correct gradient
    (grad_output * mask + scatter(grad_topk_values)); set OPTMaxK.reference_compat = True to
    reproduce the reference bit for bit.
    """
    reference_compat = False

    @staticmethod
    def forward(ctx, input, k=1):
        _require_cuda_2d(input, "OPTMaxK")
        r = _k.topk_cbsr(input, int(k), order=_k.ORDER_BANKED, want_masked=True, want_i64=True)
        ctx.save_for_backward(r["sel"])
        ctx.mark_non_differentiable(r["i64"])
        return r["masked"], r["values"], r["i64"]

    @staticmethod
    def backward(ctx, grad_output, grad_topk_values, grad_topk_indices):
        (sel,) = ctx.saved_tensors
        add = None
        if grad_topk_values is not None and not OPTMaxK.reference_compat:
            add = grad_topk_values.contiguous()
        return _k.mask_apply(grad_output.contiguous(), sel, add), None

