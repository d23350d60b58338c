"""MaxK SpGEMM autograd operator, generation 1 surface.

Same names, argument order and return arity as the reference's maxk_spgemm_function.py
(MaxKSpGEMMFunction :20-186, maxk_spgemm :188-212, MaxKSpmmWrapper :214-267), running on the
hand-written sm_100a kernels of this package:

    forward : top-k -> CBSR (one fused kernel), SpGEMM with the /in_degrees fused in the epilogue
    backward: SSpMM with the /out_degrees fused in the row stage, scatter to dense [N, 256]

Deliberate differences: the reference's arity bug (3 tensors saved, 7
unpacked, :66 vs :144; 7 Nones returned for 11 inputs, :184) is fixed; nothing is printed; there
is no cuSPARSE / torch.sparse fallback -- a failure raises.
"""
import torch
from torch.autograd import Function

import maxk_cuda_kernels

MAXK_KERNELS_AVAILABLE = True   # importing maxk_cuda_kernels raises if the library is missing


def _row_ranges(warp4_metadata, num_warps, graph_indptr, n_rows):
    """(row_begin, row_end, row plan) from the CSR indptr when given, else from the warp4 quads."""
    return maxk_cuda_kernels.rows_and_plan(warp4_metadata, num_warps, graph_indptr, n_rows)


class MaxKSpGEMMFunction(Function):
    @staticmethod
    def forward(ctx, graph_indices, graph_values, input_features, k_value,
                warp4_metadata, num_warps, graph_indptr=None, in_degrees=None, out_degrees=None,
                graph_indices_T=None, graph_values_T=None):
        n, d = input_features.shape
        k_value = int(k_value)
        if k_value < d:                                        # maxk_spgemm_function.py:51-57
            r = maxk_cuda_kernels.topk_cbsr(input_features, k_value, order=maxk_cuda_kernels.ORDER_BANKED)
            sparse_data, sparse_selector = r["values"], r["sel"]
        else:                                                  # :58-63, k >= D keeps every feature
            if d > maxk_cuda_kernels.FULL_DIM:
                raise RuntimeError("feature dim %d > 256 cannot be addressed by uint8 selectors" % d)
            sparse_data = input_features.contiguous()
            sparse_selector = torch.arange(d, device=input_features.device, dtype=torch.uint8).repeat(n, 1)
        row_begin, row_end, plan = _row_ranges(warp4_metadata, num_warps, graph_indptr, n)
        bwd_indices = graph_indices_T if graph_indices_T is not None else graph_indices
        bwd_values = graph_values_T if graph_values_T is not None else graph_values
        saved_deg = out_degrees if out_degrees is not None else torch.empty(0, device=input_features.device)
        ctx.save_for_backward(bwd_indices, bwd_values, sparse_selector, row_begin, row_end, saved_deg)
        ctx.has_out_degrees = out_degrees is not None
        ctx.input_shape = (n, d)
        return maxk_cuda_kernels.spgemm_forward_csr(
            row_begin, row_end, graph_indices, graph_values, sparse_data, sparse_selector,
            out_dim=maxk_cuda_kernels.FULL_DIM, row_div=in_degrees, plan=plan)          # :76-86

    @staticmethod
    def backward(ctx, grad_output):
        bwd_indices, bwd_values, sparse_selector, row_begin, row_end, out_degrees = ctx.saved_tensors
        grad_sparse = maxk_cuda_kernels.sspmm_backward_csr(
            row_begin, row_end, bwd_indices, bwd_values, grad_output.contiguous(), sparse_selector,
            row_div=out_degrees if ctx.has_out_degrees else None)            # :155-172
        grad_input = maxk_cuda_kernels.cbsr_scatter(grad_sparse, sparse_selector, dim=ctx.input_shape[1])  # :152,175
        return None, None, grad_input, None, None, None, None, None, None, None, None


def maxk_spgemm(graph_indices, graph_values, input_features, k_value,
                warp4_metadata=None, num_warps=0, graph_indptr=None, in_degrees=None, out_degrees=None,
                graph_indices_T=None, graph_values_T=None):
    return MaxKSpGEMMFunction.apply(
        graph_indices, graph_values, input_features, k_value,
        warp4_metadata, num_warps, graph_indptr, in_degrees, out_degrees, graph_indices_T, graph_values_T)


class MaxKSpmmWrapper:
    """Holds the warp4 metadata of one graph (maxk_spgemm_function.py:214-267)."""

    def __init__(self, graph_name="", num_warps=12, warp_max_nz=64):
        self.graph_name = graph_name
        self.warp4_metadata = None
        self.num_warps = 0
        self.num_warps_config = num_warps
        self.warp_max_nz = warp_max_nz

    def load_metadata(self, graph_name=None):
        """Reads kernels/w12_nz64_warp_4/<graph>.warp4 like the reference; raises if it is missing."""
        if graph_name is None:
            graph_name = self.graph_name
        self.warp4_metadata = maxk_cuda_kernels.load_warp4_metadata(graph_name, self.num_warps_config, self.warp_max_nz)
        self.num_warps = self.warp4_metadata.size(0) // 4
        return True

    def build_metadata(self, graph_indptr):
        """Additive: builds the same quads on the GPU from the CSR indptr (no file, no host loop)."""
        self.warp4_metadata, self.num_warps = maxk_cuda_kernels.build_warp4(graph_indptr, self.warp_max_nz)
        return True

    def spmm(self, graph_indices, graph_values, input_features, k_value, graph_indptr=None, in_degrees=None,
             out_degrees=None, graph_indices_T=None, graph_values_T=None):
        return maxk_spgemm(graph_indices, graph_values, input_features, k_value, self.warp4_metadata, self.num_warps,
                           graph_indptr, in_degrees, out_degrees, graph_indices_T, graph_values_T)
