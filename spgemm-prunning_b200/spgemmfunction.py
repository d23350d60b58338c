"""MaxK SpGEMM autograd operator, "optimized" surface (pre-computed top-k, CSC arrays for the backward).

Same names, argument order and return arity as the reference's spgemmfunction.py
(OptimizedMaxKSpGEMMFunction :18-108, optimized_maxk_spgemm :110-136, OptimizedMaxKSpmmWrapper :138-190):
the caller runs OPTMaxK once and passes (topk_values, topk_indices); the forward divides by in_degrees
after the SpGEMM and the backward divides grad_output by out_degrees before the SSpMM -- both fused into
the kernels here.  Like the reference (:97-105) the backward runs on the CSC arrays with the *same*
(CSR) quads, which is only meaningful on the undirected graphs it trains on.
"""
import torch
from torch.autograd import Function

import maxk_cuda_kernels
from maxk_spgemm_function import _row_ranges

MAXK_KERNELS_AVAILABLE = True


class OptimizedMaxKSpGEMMFunction(Function):
    @staticmethod
    def forward(ctx, graph_indices, graph_values, topk_values, topk_indices,
                warp4_metadata, num_warps, graph_indptr, in_degrees, out_degrees,
                graph_indices_T, graph_values_T):
        for name, v in (("warp4_metadata", warp4_metadata), ("topk_values", topk_values),
                        ("topk_indices", topk_indices), ("in_degrees", in_degrees), ("out_degrees", out_degrees),
                        ("graph_indices_T", graph_indices_T), ("graph_values_T", graph_values_T)):
            if v is None:
                raise RuntimeError("%s REQUIRED" % name)                              # spgemmfunction.py:45-48
        sparse_selector = topk_indices if topk_indices.dtype == torch.uint8 else topk_indices.to(torch.uint8)  # :51
        row_begin, row_end, plan = _row_ranges(warp4_metadata, num_warps, graph_indptr, topk_values.size(0))
        ctx.save_for_backward(graph_indices_T, graph_values_T, sparse_selector, out_degrees, row_begin, row_end)
        return maxk_cuda_kernels.spgemm_forward_csr(
            row_begin, row_end, graph_indices, graph_values, topk_values, sparse_selector,
            out_dim=maxk_cuda_kernels.FULL_DIM, row_div=in_degrees, plan=plan)                   # :64-77

    @staticmethod
    def backward(ctx, grad_output):
        graph_indices_T, graph_values_T, sparse_selector, out_degrees, row_begin, row_end = ctx.saved_tensors
        grad_sparse = maxk_cuda_kernels.sspmm_backward_csr(
            row_begin, row_end, graph_indices_T, graph_values_T, grad_output.contiguous(), sparse_selector,
            row_div=out_degrees)                                                      # :94-105
        return (None, None, grad_sparse, None, None, None, None, None, None, None, None)


def optimized_maxk_spgemm(graph_indices, graph_values, topk_values, topk_indices,
                          warp4_metadata, num_warps, graph_indptr, in_degrees, out_degrees,
                          graph_indices_T, graph_values_T):
    return OptimizedMaxKSpGEMMFunction.apply(
        graph_indices, graph_values, topk_values, topk_indices,
        warp4_metadata, num_warps, graph_indptr, in_degrees, out_degrees,
        graph_indices_T, graph_values_T)


class OptimizedMaxKSpmmWrapper:
    def __init__(self, graph_name="", num_warps=12, warp_max_nz=64):
        self.graph_name = graph_name
        self.warp4_metadata = None
        self.num_warps = 0
        self.num_warps_config = num_warps
        self.warp_max_nz = warp_max_nz

    def load_metadata(self, graph_name=None):
        if graph_name is None:
            graph_name = self.graph_name
        if not graph_name:
            raise RuntimeError("graph_name REQUIRED")                                 # :154
        self.warp4_metadata = maxk_cuda_kernels.load_warp4_metadata(graph_name, self.num_warps_config, self.warp_max_nz)
        self.num_warps = self.warp4_metadata.size(0) // 4
        return True

    def build_metadata(self, graph_indptr):
        self.warp4_metadata, self.num_warps = maxk_cuda_kernels.build_warp4(graph_indptr, self.warp_max_nz)
        return True

    def spmm(self, graph_indices, graph_values, topk_values, topk_indices,
             graph_indptr, in_degrees, out_degrees, graph_indices_T, graph_values_T):
        if self.warp4_metadata is None:
            raise RuntimeError("Metadata not loaded")                                 # :186
        return optimized_maxk_spgemm(graph_indices, graph_values, topk_values, topk_indices,
                                     self.warp4_metadata, self.num_warps, graph_indptr,
                                     in_degrees, out_degrees, graph_indices_T, graph_values_T)
