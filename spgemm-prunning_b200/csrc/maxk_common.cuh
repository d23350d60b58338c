// maxk_common.cuh -- shared device helpers for the sm_100a MaxK kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/maxk_b200.h"

namespace maxk {

constexpr int kAccDim = 256;          // accumulator width: uint8 selectors address at most 256 columns
constexpr int kNumSMsB200 = 148;
constexpr int kLongRow = 4096;        // rows with more edges are handled by a whole CTA
constexpr unsigned kFullMask = 0xffffffffu;

// Workspace layout shared by forward and backward (see maxk_spgemm_workspace_bytes).
struct SchedWorkspace {
    int row_counter;       // dynamic row scheduler
    int long_count;        // number of long rows appended by the main kernel
    int long_counter;      // scheduler of the long-row kernel
    int pad;
    // followed by int long_rows[n_rows]
};

// ---------------------------------------------------------------------------------------------
// Banked lane layout shared by the forward accumulator and the backward staged gradient row.
//
// The SM's LSU data pipe moves one 128-byte wavefront per clock and is the limiter of both
// kernels (ncu: l1tex__data_pipe_lsu_wavefronts 96 % / 90 % of peak in the first version), so
// the layout is chosen to minimise wavefronts per edge:
//   * a lane owns EPL consecutive CBSR entries of one edge (one 16-byte value load + one 4-byte
//     selector load for EPL = 4): L = k/EPL lanes cover an edge, EPI = 32/L edges per warp
//     instruction.  A request now touches EPI full rows instead of one (the LDG request rate,
//     not L2 bandwidth, capped the 4-byte-per-lane gather at 35 G rows/s vs 116 G rows/s).
//   * every edge slot q has its own copy of the 256 columns, and copy q only lives in banks
//     [q*L, q*L+L): word address of (column c, copy q) = (c / L) * 32 + q * L + (c % L).
//     Lanes of different edge slots can therefore never collide, and inside a slot a conflict
//     needs two of the slot's L lanes to hold columns with equal c % L.  The top-k kernel emits
//     rows sorted by (c % L, c) (MAXK_ORDER_BANKED), which spreads each residue class over the
//     EPL accumulate instructions as evenly as possible (McNaughton wrap-around).
// k = 8 uses EPL = 2; any other k falls back to L = 32, one copy, natural layout (address = c).
// ---------------------------------------------------------------------------------------------
template <int K>
struct Lay {
    static constexpr bool kFast = (K == 8 || K == 16 || K == 32 || K == 64);
    static constexpr int EPL = kFast ? (K == 8 ? 2 : 4) : 1;   // entries per lane
    static constexpr int L = kFast ? K / EPL : 32;             // lanes per edge == banks per copy
    static constexpr int EPI = 32 / L;                         // edges per warp instruction == copies
    static constexpr int kWords = (kAccDim / L) * 32;          // floats per warp (== EPI * 256)
    __device__ static __forceinline__ int word(int col) { return (col / L) * 32 + (col % L); }  // + q * L
};

// residue modulus of MAXK_ORDER_BANKED for a given k (host + device): 4 = the lanes per slot of slots.cuh
// for the k that have a vectorised path, 1 (plain column order) otherwise
__host__ __device__ inline int banked_modulus(int k)
{
    return (k == 8 || k == 16 || k == 32 || k == 64 || k == 96 || k == 128) ? 4 : 1;
}

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }

// L2 residency hints.  The streams (CSR indices/values, dense feature / gradient rows) are read once
// and are far larger than L2, the gathered / reduced operands (CBSR values + selectors, the sampled
// gradient) are small and re-used ~degree times: streams are loaded "evict first" so that they do
// not push the re-used lines out of the 126 MB L2 (measured on the Yelp shape: DRAM traffic of the
// backward was 2.4x the algorithmic bytes without the hints).
__device__ __forceinline__ uint64_t policy_evict_first()
{
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t policy_evict_last()
{
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}

// Streaming (read-once) loads: no L1 allocation, first to leave L2.
__device__ __forceinline__ int ld_stream_i32(const int *p)
{
    int v;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.s32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(policy_evict_first()));
    return v;
}
__device__ __forceinline__ float ld_stream_f32(const float *p)
{
    float v;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(v) : "l"(p), "l"(policy_evict_first()));
    return v;
}
__device__ __forceinline__ float4 ld_stream_f32x4(const float *p)
{
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p), "l"(policy_evict_first()));
    return v;
}
// 32 bytes per lane in one instruction (sm_100: LDG.E.256), 32-byte aligned address.
__device__ __forceinline__ void ld_stream_f32x8(const float *p, float (&v)[8])
{
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8], %9;"
                 : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7])
                 : "l"(p), "l"(policy_evict_first()));
}
// Re-used (gathered) operands: keep them in L2.
__device__ __forceinline__ float4 ld_keep_f32x4(const float *p, uint64_t pol)
{
    float4 v;
    asm volatile("ld.global.nc.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ float2 ld_keep_f32x2(const float *p, uint64_t pol)
{
    float2 v;
    asm volatile("ld.global.nc.L2::cache_hint.v2.f32 {%0,%1}, [%2], %3;" : "=f"(v.x), "=f"(v.y) : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ uint32_t ld_keep_u32(const void *p, uint64_t pol)
{
    uint32_t v;
    asm volatile("ld.global.nc.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ uint32_t ld_keep_u16(const void *p, uint64_t pol)
{
    unsigned short v;
    asm volatile("ld.global.nc.L2::cache_hint.u16 %0, [%1], %2;" : "=h"(v) : "l"(p), "l"(pol));
    return v;
}
// Streaming stores for outputs that are not re-read by this kernel.
__device__ __forceinline__ void st_stream_f32(float *p, float v)
{
    asm volatile("st.global.cs.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}
__device__ __forceinline__ void st_stream_f32x4(float *p, float4 v)
{
    asm volatile("st.global.cs.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void st_stream_f32x8(float *p, const float (&v)[8])
{
    asm volatile("st.global.cs.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]),
                 "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]) : "memory");
}
// 16-byte vector reduction into global memory (sm_90+): one L2 atomic per 4 floats.
__device__ __forceinline__ void red_add_f32x4(float *p, float a, float b, float c, float d, uint64_t pol)
{
    asm volatile("red.global.add.L2::cache_hint.v4.f32 [%0], {%1,%2,%3,%4}, %5;" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d), "l"(pol) : "memory");
}
__device__ __forceinline__ void red_add_f32x2(float *p, float a, float b, uint64_t pol)
{
    asm volatile("red.global.add.L2::cache_hint.v2.f32 [%0], {%1,%2}, %3;" ::"l"(p), "f"(a), "f"(b), "l"(pol) : "memory");
}

// x / d with one reciprocal per row instead of one IEEE division per element: q = x*r followed by one
// residual correction (Markstein) is correctly rounded except in rare double-rounding cases (<= 1 ulp,
// far inside the 1e-5 parity tolerance) and costs 3 FMA-pipe instructions instead of a ~30-instruction
// division with a slow-path call (the divisions were 23 % of the instructions of a low-degree row).
__device__ __forceinline__ float div_by_recip(float x, float d, float r)
{
    const float q = x * r;
    return fmaf(fmaf(-q, d, x), r, q);
}

// True when 1/d is a normal number, i.e. div_by_recip is usable; zero / infinite / NaN divisors (and
// divisors so large that the reciprocal is denormal-flushed) take the plain IEEE division instead.
__device__ __forceinline__ bool recip_usable(float r) { return r != 0.f && fabsf(r) <= 3.0e38f; }
__device__ __forceinline__ float div_guarded(float x, float d)
{
    const float r = 1.0f / d;
    return recip_usable(r) ? div_by_recip(x, d, r) : x / d;
}

inline int status_from_cuda(cudaError_t e) { return e == cudaSuccess ? MAXK_OK : (int)e; }

inline int device_sm_count()
{
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    return sms > 0 ? sms : kNumSMsB200;
}

// Launch configuration of a (main, long-row) kernel pair, cached per device: the opt-in dynamic shared
// memory size is a per-device function attribute, so a process that drives several GPUs has to set it on
// each of them (one process per GPU is the normal deployment, but nothing here relies on it).
constexpr int kMaxCachedDevices = 64;
struct LaunchConfig {
    bool configured;
    int blocks_per_sm;
    int sms;
};
template <typename MainKernel, typename LongKernel>
inline cudaError_t configure_pair(LaunchConfig &c, MainKernel main_kernel, LongKernel long_kernel, int main_threads,
                                  size_t smem_main, size_t smem_long)
{
    cudaError_t err = cudaFuncSetAttribute(main_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_main);
    if (err != cudaSuccess) return err;
    err = cudaFuncSetAttribute(long_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_long);
    if (err != cudaSuccess) return err;
    int blocks = 0;
    err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks, main_kernel, main_threads, smem_main);
    if (err != cudaSuccess) return err;
    c.blocks_per_sm = blocks < 1 ? 1 : blocks;
    c.sms = device_sm_count();
    c.configured = true;
    return cudaSuccess;
}

}  // namespace maxk
