// maxk_common.cuh -- shared device helpers for the sm_100a MaxK kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/maxk_b200.h"

namespace maxk {

constexpr int kAccDim = 256;          // accumulator width: uint8 selectors address at most 256 columns
constexpr int kNumSMsB200 = 148;
constexpr int kLongRow = 4096;        // rows with more edges are handled by a whole CTA

// Workspace layout shared by forward and backward (see maxk_spgemm_workspace_bytes).
struct SchedWorkspace {
    int row_counter;       // dynamic row scheduler
    int long_count;        // number of long rows appended by the main kernel
    int long_counter;      // scheduler of the long-row kernel
    int pad;
    // followed by int long_rows[n_rows]
};

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }

// Streaming (read-once) loads: CSR indices / values are touched exactly once per pass,
// keep them out of L1 so the gathered CBSR rows own the cache.
__device__ __forceinline__ int ld_stream_i32(const int *p)
{
    int v;
    asm volatile("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ float ld_stream_f32(const float *p)
{
    float v;
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ float4 ld_stream_f32x4(const float *p)
{
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}
// Streaming stores for outputs that are not re-read by this kernel.
__device__ __forceinline__ void st_stream_f32x4(float *p, float4 v)
{
    asm volatile("st.global.cs.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
// 16-byte vector reduction into global memory (sm_90+): one L2 atomic per 4 floats.
__device__ __forceinline__ void red_add_f32x4(float *p, float a, float b, float c, float d)
{
    asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void red_add_f32x2(float *p, float a, float b)
{
    asm volatile("red.global.add.v2.f32 [%0], {%1,%2};" ::"l"(p), "f"(a), "f"(b) : "memory");
}

// Order-preserving key: larger key <=> larger value; NaN largest; -0 == +0.
__device__ __forceinline__ uint32_t order_key(float f)
{
    uint32_t b = __float_as_uint(f);
    if ((b & 0x7fffffffu) > 0x7f800000u) return 0xffffffffu;
    if (b == 0x80000000u) b = 0u;
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

inline int status_from_cuda(cudaError_t e) { return e == cudaSuccess ? MAXK_OK : (int)e; }

}  // namespace maxk
