"""MaxK SpGEMM autograd operator, generation 4 surface (pre-computed top-k, undirected graphs).

Same names and signatures as the reference's `spgemmfunction_v4` (MaxKSpGEMMFunction :19-101,
maxk_spgemm :103-124, MaxKSpmmWrapper :126-174): the caller runs OPTMaxK once and passes
(topk_values, topk_indices); forward returns the degree-normalised [N, 256] aggregate, backward
returns the gradient w.r.t. topk_values [N, k].  Both degree divisions are fused into the kernels.
"""
import torch
from torch.autograd import Function

import maxk_cuda_kernels
from maxk_spgemm_function import _row_ranges

MAXK_KERNELS_AVAILABLE = True


class MaxKSpGEMMFunction(Function):
    @staticmethod
    def forward(ctx, graph_indices, graph_values, topk_values, topk_indices,
                warp4_metadata, num_warps, graph_indptr, degrees):
        if topk_values is None or topk_indices is None:
            raise RuntimeError("topk_values / topk_indices REQUIRED")            # spgemmfunction_v4:45-47
        if degrees is None:
            raise RuntimeError("degrees REQUIRED for normalization")              # :48
        sparse_selector = topk_indices if topk_indices.dtype == torch.uint8 else topk_indices.to(torch.uint8)  # :51
        row_begin, row_end, plan = _row_ranges(warp4_metadata, num_warps, graph_indptr, topk_values.size(0))
        ctx.save_for_backward(graph_indices, graph_values, sparse_selector, degrees, row_begin, row_end)
        return maxk_cuda_kernels.spgemm_forward_csr(
            row_begin, row_end, graph_indices, graph_values, topk_values, sparse_selector,
            out_dim=maxk_cuda_kernels.FULL_DIM, row_div=degrees, plan=plan)                  # :61-72

    @staticmethod
    def backward(ctx, grad_output):
        graph_indices, graph_values, sparse_selector, degrees, row_begin, row_end = ctx.saved_tensors
        grad_sparse = maxk_cuda_kernels.sspmm_backward_csr(
            row_begin, row_end, graph_indices, graph_values, grad_output.contiguous(), sparse_selector,
            row_div=degrees)                                                      # :87-98
        return None, None, grad_sparse, None, None, None, None, None


def maxk_spgemm(graph_indices, graph_values, topk_values, topk_indices,
                warp4_metadata, num_warps, graph_indptr, degrees):
    return MaxKSpGEMMFunction.apply(graph_indices, graph_values, topk_values, topk_indices,
                                    warp4_metadata, num_warps, graph_indptr, degrees)


class MaxKSpmmWrapper:
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
            raise RuntimeError("graph_name REQUIRED")
        self.warp4_metadata = maxk_cuda_kernels.load_warp4_metadata(graph_name, self.num_warps_config, self.warp_max_nz)
        self.num_warps = self.warp4_metadata.size(0) // 4
        return True

    def build_metadata(self, graph_indptr):
        self.warp4_metadata, self.num_warps = maxk_cuda_kernels.build_warp4(graph_indptr, self.warp_max_nz)
        return True

    def spmm(self, graph_indices, graph_values, topk_values, topk_indices, graph_indptr, degrees):
        if self.warp4_metadata is None and graph_indptr is None:
            raise RuntimeError("Metadata not loaded")
        return maxk_spgemm(graph_indices, graph_values, topk_values, topk_indices,
                           self.warp4_metadata, self.num_warps, graph_indptr, degrees)
