// ref_harness_fwd.cu -- TEST INFRASTRUCTURE ONLY (see maxk_oracle.c header).
//
// Compiles the reference's forward kernel UNMODIFIED, from where it lies
// (-I/root/reference/kernels; no source is copied into this repo), and exposes a
// plain C launcher with the launch geometry of the reference binding
// (cuda_kernel_bindings.cpp:71 zero-fill, :77-85 grid/block/shared size).
// Output goes to oracle/_ref/libmaxk_ref.so (git-ignored, travels to the GPU box).
#include <cstdint>
#include <cuda_runtime.h>
#include "spmm_maxk.cu"   // /root/reference/kernels/spmm_maxk.cu

std::string base_dir, graph;   // the reference declares these extern (spmm_maxk.cu:9)

extern "C" int ref_spmm_maxk_forward(const int *warp4, const int *idx, const float *val,
                                     const float *data, const uint8_t *sel, float *out,
                                     int num_v, int num_e, int feat_in, int dim_sparse,
                                     int num_warps, void *stream)
{
    cudaStream_t s = (cudaStream_t)stream;
    cudaMemsetAsync(out, 0, sizeof(float) * (size_t)num_v * feat_in, s);
    if (num_warps <= 0) return (int)cudaGetLastError();
    int block_num = (num_warps + WARPS_PER_BLOCK - 1) / WARPS_PER_BLOCK;
    int shared = WARPS_PER_BLOCK * feat_in * (int)sizeof(float);
    spmm_kernel_opt2_sparse_v3<<<block_num, WARPS_PER_BLOCK * EXT_WARP_DIM, shared, s>>>(
        warp4, idx, val, data, sel, out, num_v, num_e, feat_in, dim_sparse, num_warps);
    return (int)cudaGetLastError();
}
