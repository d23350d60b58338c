"""MaxK SpGEMM autograd operator, generation 3 surface (CSR forward, CSC + CSC-warp4 backward).

Same names, argument order and return arity as the reference's spgemmfunction_v3.py
(MaxKSpGEMMFunction :22-156, maxk_spgemm :158-189, MaxKSpmmWrapper :191-271): the forward runs on the CSR
arrays with the CSR quads, the backward hands the CSC arrays (graph_indices_T / graph_values_T) and the
CSC quads to the SSpMM kernel.  The kernel evaluates SURVEY a-6's formula on whatever arrays arrive, so
with a true transpose this is sample(A g), which equals the adjoint sample(A^T g) on the undirected
graphs the reference trains on (dataset_gen.py:44-58); pass the CSR arrays twice for the exact adjoint
of a directed graph.

Deliberate difference: the reference multiplies grad_output by the *input's* top-k mask before the
backward kernel (:118-119), which is not the gradient of its forward.  It is applied only when
`MaxKSpGEMMFunction.reference_compat` is set.  Nothing is printed, nothing falls back.
"""
import torch
from torch.autograd import Function

import maxk_cuda_kernels

MAXK_KERNELS_AVAILABLE = True


def _rows(warp4_metadata, num_warps, n_rows, what):
    if warp4_metadata is None:
        raise RuntimeError("%s metadata required" % what)                            # spgemmfunction_v3.py:57-58
    return maxk_cuda_kernels.rows_and_plan(warp4_metadata, num_warps, None, n_rows)


class MaxKSpGEMMFunction(Function):
    reference_compat = False

    @staticmethod
    def forward(ctx, graph_indices, graph_values, input_features, k_value,
                warp4_metadata_csr, num_warps_csr, graph_indptr,
                in_degrees, out_degrees,
                graph_indices_T, graph_values_T,
                warp4_metadata_csc, num_warps_csc):
        n, d = input_features.shape
        k_value = int(k_value)
        if graph_indices_T is None or graph_values_T is None:
            raise RuntimeError("graph_indices_T / graph_values_T (CSC arrays) required for the backward pass")
        if k_value < d:                                                               # :61-64
            r = maxk_cuda_kernels.topk_cbsr(input_features, k_value, order=maxk_cuda_kernels.ORDER_BANKED)
            sparse_data, sparse_selector = r["values"], r["sel"]
        else:                                                                         # :65-71
            if d > maxk_cuda_kernels.FULL_DIM:
                raise RuntimeError("feature dim %d > 256 cannot be addressed by uint8 selectors" % d)
            sparse_data = input_features.contiguous()
            sparse_selector = torch.arange(d, device=input_features.device, dtype=torch.uint8).repeat(n, 1)
        if graph_indptr is not None:
            if warp4_metadata_csr is None:
                raise RuntimeError("CSR metadata required")
            row_begin, row_end, plan = maxk_cuda_kernels.plan_for_indptr(graph_indptr)
        else:
            row_begin, row_end, plan = _rows(warp4_metadata_csr, num_warps_csr, n, "CSR")
        t_begin, t_end, _ = _rows(warp4_metadata_csc, num_warps_csc, n, "CSC")
        saved_deg = out_degrees if out_degrees is not None else torch.empty(0, device=input_features.device)
        ctx.save_for_backward(graph_indices_T, graph_values_T, sparse_selector, t_begin, t_end, saved_deg)
        ctx.has_out_degrees = out_degrees is not None
        ctx.input_shape = (n, d)
        return maxk_cuda_kernels.spgemm_forward_csr(
            row_begin, row_end, graph_indices, graph_values, sparse_data, sparse_selector,
            out_dim=maxk_cuda_kernels.FULL_DIM, row_div=in_degrees, plan=plan)                   # :85-99

    @staticmethod
    def backward(ctx, grad_output):
        graph_indices_T, graph_values_T, sparse_selector, t_begin, t_end, out_degrees = ctx.saved_tensors
        grad_output = grad_output.contiguous()
        if MaxKSpGEMMFunction.reference_compat:                                       # :118-119
            grad_output = maxk_cuda_kernels.mask_apply(grad_output, sparse_selector)
        grad_sparse = maxk_cuda_kernels.sspmm_backward_csr(
            t_begin, t_end, graph_indices_T, graph_values_T, grad_output, sparse_selector,
            row_div=out_degrees if ctx.has_out_degrees else None)                     # :121-143
        grad_input = maxk_cuda_kernels.cbsr_scatter(grad_sparse, sparse_selector, dim=ctx.input_shape[1])  # :146-147
        return (None, None, grad_input, None, None, None, None, None, None, None, None, None, None)


def maxk_spgemm(graph_indices, graph_values, input_features, k_value,
                warp4_metadata_csr, num_warps_csr, graph_indptr,
                in_degrees, out_degrees,
                graph_indices_T, graph_values_T,
                warp4_metadata_csc, num_warps_csc):
    return MaxKSpGEMMFunction.apply(
        graph_indices, graph_values, input_features, k_value,
        warp4_metadata_csr, num_warps_csr, graph_indptr,
        in_degrees, out_degrees, graph_indices_T, graph_values_T,
        warp4_metadata_csc, num_warps_csc)


class MaxKSpmmWrapper:
    """Holds the CSR and the CSC quads of one graph (spgemmfunction_v3.py:191-271)."""

    def __init__(self, graph_name="", num_warps=12, warp_max_nz=64):
        self.graph_name = graph_name
        self.warp4_metadata_csr = None
        self.num_warps_csr = 0
        self.warp4_metadata_csc = None
        self.num_warps_csc = 0
        self.num_warps_config = num_warps
        self.warp_max_nz = warp_max_nz

    def load_metadata(self, graph_name=None):
        """Reads <g>.warp4 and <g>.warp4_csc; raises RuntimeError if either is missing (:222-239)."""
        if graph_name is None:
            graph_name = self.graph_name
        try:
            self.warp4_metadata_csr = maxk_cuda_kernels.load_warp4_metadata(graph_name, self.num_warps_config, self.warp_max_nz)
        except Exception as e:
            raise RuntimeError("Failed to load CSR metadata for %s: %s" % (graph_name, e))
        self.num_warps_csr = self.warp4_metadata_csr.size(0) // 4
        try:
            self.warp4_metadata_csc = maxk_cuda_kernels.load_warp4_metadata_csc(graph_name, self.num_warps_config, self.warp_max_nz)
        except Exception as e:
            raise RuntimeError("Failed to load CSC metadata for %s: %s. Run graph_loader.generate_meta first."
                               % (graph_name, e))
        self.num_warps_csc = self.warp4_metadata_csc.size(0) // 4
        return True

    def build_metadata(self, graph_indptr, graph_indptr_T):
        """Additive: both sets of quads from the CSR / CSC indptr on the GPU (no files)."""
        self.warp4_metadata_csr, self.num_warps_csr = maxk_cuda_kernels.build_warp4(graph_indptr, self.warp_max_nz)
        self.warp4_metadata_csc, self.num_warps_csc = maxk_cuda_kernels.build_warp4(graph_indptr_T, self.warp_max_nz)
        return True

    def spmm(self, graph_indices, graph_values, input_features, k_value,
             graph_indptr, in_degrees, out_degrees, graph_indices_T, graph_values_T):
        if self.warp4_metadata_csr is None:
            raise RuntimeError("CSR metadata not loaded")
        if self.warp4_metadata_csc is None:
            raise RuntimeError("CSC metadata not loaded")
        return maxk_spgemm(graph_indices, graph_values, input_features, k_value,
                           self.warp4_metadata_csr, self.num_warps_csr, graph_indptr,
                           in_degrees, out_degrees, graph_indices_T, graph_values_T,
                           self.warp4_metadata_csc, self.num_warps_csc)
