"""DirectMaxKKernels -- bench / validation facade, same surface as the reference's
direct_kernel_interface.py (:24-460): load_warp4_metadata, generate_maxk_sparse_data,
run_forward_kernel, run_backward_kernel, validate_against_cusparse, benchmark_all_k_values.

Differences: nothing falls back (a failing kernel raises), per-call emoji prints are gone (the
benchmark keeps the `main.cu` result lines), and build_warp4_metadata() builds the quads on the
GPU when no .warp4 file exists.
"""
import numpy as np
import torch

import maxk_cuda_kernels

DIRECT_KERNELS_AVAILABLE = True


class DirectMaxKKernels:
    def __init__(self, graph_name=""):
        self.graph_name = graph_name
        self.warp4_metadata = None
        self.num_warps = 0

    # -- metadata ------------------------------------------------------------------------------
    def load_warp4_metadata(self, graph_name=None, num_warps=12, warp_max_nz=64):
        """direct_kernel_interface.py:35-56; returns False when the file is missing, like the reference."""
        if graph_name is None:
            graph_name = self.graph_name
        try:
            self.warp4_metadata = maxk_cuda_kernels.load_warp4_metadata(graph_name, num_warps, warp_max_nz)
        except RuntimeError:
            return False
        self.num_warps = self.warp4_metadata.size(0) // 4
        return True

    def build_warp4_metadata(self, graph_data, warp_max_nz=64):
        """Additive: the same quads from graph_data['indptr'] on the GPU (replaces running generate_meta.py)."""
        self.warp4_metadata, self.num_warps = maxk_cuda_kernels.build_warp4(graph_data["indptr"], warp_max_nz)
        return True

    def _require_metadata(self):
        if self.warp4_metadata is None:
            raise RuntimeError("Warp4 metadata not loaded. Call load_warp4_metadata() first")

    # -- CBSR ----------------------------------------------------------------------------------
    def generate_maxk_sparse_data(self, input_features, dim_k, use_cuda_topk=True):
        """direct_kernel_interface.py:58-91 -> (sparse_data fp32 [N,k], sparse_selector uint8 [N,k]).

        use_cuda_topk=True: our exact top-k kernel; False: torch.topk, as in the reference."""
        if use_cuda_topk:
            r = maxk_cuda_kernels.topk_cbsr(input_features, dim_k, order=maxk_cuda_kernels.ORDER_VALUE_DESC)
            return r["values"], r["sel"]
        topk_values, topk_indices = torch.topk(input_features, dim_k, dim=1)
        return topk_values, topk_indices.to(torch.uint8)

    # -- kernels -------------------------------------------------------------------------------
    def run_forward_kernel(self, graph_data, input_features, dim_k, timing=True, use_cuda_topk=True):
        """direct_kernel_interface.py:93-159 -> (output [N,256], avg ms)."""
        self._require_metadata()
        sparse_data, sparse_selector = self.generate_maxk_sparse_data(input_features, dim_k, use_cuda_topk)
        avg_time = 0.0
        if timing:
            times = maxk_cuda_kernels.benchmark_spmm_maxk(
                self.warp4_metadata, graph_data["indices"], graph_data["values"], sparse_data, sparse_selector,
                self.num_warps, dim_k, num_runs=4)
            avg_time = float(np.mean(times))
        output = maxk_cuda_kernels.spmm_maxk_forward(
            self.warp4_metadata, graph_data["indices"], graph_data["values"], sparse_data, sparse_selector,
            self.num_warps, dim_k)
        return output, avg_time

    def run_backward_kernel(self, graph_data, grad_output, dim_k, timing=True, use_cuda_topk=True):
        """direct_kernel_interface.py:161-219 -> (grad_input [N,k], avg ms); selector = top-k of grad_output."""
        self._require_metadata()
        _, sparse_selector = self.generate_maxk_sparse_data(grad_output, dim_k, use_cuda_topk)
        args = (self.warp4_metadata, graph_data["indices"], graph_data["values"], grad_output, sparse_selector,
                self.num_warps, dim_k)
        if not timing:
            return maxk_cuda_kernels.spmm_maxk_backward(*args), 0.0
        timer, times = maxk_cuda_kernels.CudaTimer(), []
        for i in range(8):                       # 4 warm-up + 4 timed, :190-206
            timer.start()
            grad_input = maxk_cuda_kernels.spmm_maxk_backward(*args)
            elapsed = timer.stop()
            if i >= 4:
                times.append(elapsed)
        return grad_input, float(np.mean(times))

    # -- validation ----------------------------------------------------------------------------
    def validate_against_cusparse(self, graph_data, input_features, dim_k, tolerance=0.001, use_cuda_topk=True):
        """direct_kernel_interface.py:221-372: same top-k fed to the MaxK kernel and to a dense SpMM of the
        scattered input; max |diff| at the positions where the input was non-zero must be < tolerance."""
        self._require_metadata()
        sparse_data, sparse_selector = self.generate_maxk_sparse_data(input_features, dim_k, use_cuda_topk)
        sparse_input = maxk_cuda_kernels.cbsr_scatter(sparse_data, sparse_selector, dim=input_features.size(1))
        maxk_output = maxk_cuda_kernels.spmm_maxk_forward(
            self.warp4_metadata, graph_data["indices"], graph_data["values"], sparse_data, sparse_selector,
            self.num_warps, dim_k)
        dense_output = maxk_cuda_kernels.cusparse_spmm(graph_data["indptr"], graph_data["indices"],
                                                       graph_data["values"], sparse_input, timing=False)
        if maxk_output.shape != dense_output.shape:
            return False
        mask = sparse_input != 0
        diff = (maxk_output - dense_output).abs()[mask]
        self.last_validation = {"max_error": float(diff.max()) if diff.numel() else 0.0,
                                "avg_error": float(diff.mean()) if diff.numel() else 0.0}
        return self.last_validation["max_error"] < tolerance

    # -- benchmark -----------------------------------------------------------------------------
    def benchmark_all_k_values(self, graph_data, dim_origin=256, k_values=(16, 32, 64), num_runs=4, use_cuda_topk=True,
                               verbose=True):
        """direct_kernel_interface.py:374-460 (the loop of kernels/main.cu:109-169)."""
        self._require_metadata()
        v_num = graph_data["indptr"].size(0) - 1
        torch.manual_seed(123)
        input_features = torch.rand(v_num, dim_origin, device="cuda", dtype=torch.float32)
        results = {}
        if verbose:
            print("num graph dim_origin dim_k kernel time(ms)")
        for dim_k in k_values:
            if dim_k > 64:                       # :428, kernels/main.cu:114
                continue
            _, t_fwd = self.run_forward_kernel(graph_data, input_features, dim_k, True, use_cuda_topk)
            grad_output = torch.rand_like(input_features)
            _, t_bwd = self.run_backward_kernel(graph_data, grad_output, dim_k, True, use_cuda_topk)
            results[dim_k] = {"forward_time": t_fwd, "backward_time": t_bwd}
            if verbose:
                print("1/1 %s %d %d maxk %.3f" % (self.graph_name, dim_origin, dim_k, t_fwd))
                print("1/1 %s %d %d maxk_backward %.3f" % (self.graph_name, dim_origin, dim_k, t_bwd))
        return results
