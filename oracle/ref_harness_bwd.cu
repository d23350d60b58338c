// ref_harness_bwd.cu -- TEST INFRASTRUCTURE ONLY (see maxk_oracle.c header).
//
// Compiles the reference's backward kernel UNMODIFIED from /root/reference/kernels and
// exposes a C launcher with the geometry of cuda_kernel_bindings.cpp:128 (zero-fill),
// :134-142 (grid/block/shared size).
#include <cstdint>
#include <cuda_runtime.h>
#include "spmm_maxk_backward.cu"   // /root/reference/kernels/spmm_maxk_backward.cu

extern "C" int ref_spmm_maxk_backward(const int *warp4, const int *idx, const float *val,
                                      const float *grad, const uint8_t *sel, float *gs,
                                      int num_v, int num_e, int feat_in, int dim_sparse,
                                      int num_warps, void *stream)
{
    cudaStream_t s = (cudaStream_t)stream;
    cudaMemsetAsync(gs, 0, sizeof(float) * (size_t)num_v * dim_sparse, s);
    if (num_warps <= 0) return (int)cudaGetLastError();
    int block_num = (num_warps + WARPS_PER_BLOCK - 1) / WARPS_PER_BLOCK;
    int shared = WARPS_PER_BLOCK * feat_in * (int)sizeof(float);
    spmm_kernel_opt2_sparse_backward_v3<<<block_num, WARPS_PER_BLOCK * EXT_WARP_DIM, shared, s>>>(
        warp4, idx, val, grad, sel, gs, num_v, num_e, feat_in, dim_sparse, num_warps);
    return (int)cudaGetLastError();
}
