// spgemm_fwd.cu -- forward row-wise-product SpGEMM, CSR adjacency x CBSR features (sm_100a).
//
// Replaces spmm_kernel_opt2_sparse_v3 (reference kernels/spmm_maxk.cu:17-106).  Design
// (DESIGN.md "Forward SpGEMM"):
//   * one warp OWNS one output row: the whole edge list of the row is reduced by that warp
//     into a private 256-float shared-memory accumulator and the row is written exactly
//     once with 16-byte stores.  No global atomics, no output memset, fixed reduction
//     order (the reference flushes 256 global atomics per <=64-edge segment,
//     spmm_maxk.cu:101-105, into a torch::zeros output, cuda_kernel_bindings.cpp:71).
//   * CSR indices/values are read as coalesced 128-byte streaming loads, 32 edges per
//     warp instruction, and handed to the lanes by shuffles (the reference issues one
//     scalar broadcast __ldg per edge, spmm_maxk.cu:72-73).
//   * UNROLL independent neighbour-row gathers are in flight per warp before the first
//     accumulate, so the idx -> CBSR row -> accumulator chain is pipelined.
//   * k < 32 packs 32/k edges into one warp instruction, each edge slot with its own
//     accumulator copy, so all 32 lanes work for every k (the reference idles
//     1 - k/32 of its warps, spmm_maxk.cu:27,64).
//   * rows are handed out dynamically (one global counter), rows longer than kLongRow
//     are deferred to a whole-CTA kernel, so skewed graphs do not serialise on one warp.
//   * the degree normalisation the reference does in a separate PyTorch pass
//     (maxk_spgemm_function.py:86) is fused into the row epilogue.
#include "maxk_common.cuh"

namespace maxk {

constexpr int kFwdThreads = 256;
constexpr int kFwdWarps = kFwdThreads / 32;
constexpr int kLongThreads = 512;
constexpr int kLongWarps = kLongThreads / 32;
constexpr unsigned kFull = 0xffffffffu;

// ---------------------------------------------------------------------------------
// Edge accumulation for one row segment [b, e), visiting 32-edge batches
// batch0, batch0+stride, ...  K = compile-time k (8/16/32), EPI = 32/K edges per instruction.
// ---------------------------------------------------------------------------------
template <int K, int UNROLL>
__device__ __forceinline__ void accumulate_small_k(const int *__restrict__ idx, const float *__restrict__ val,
                                                   const float *__restrict__ cval,
                                                   const uint8_t *__restrict__ csel, float *acc, int b, int e,
                                                   int batch0, int stride)
{
    constexpr int EPI = 32 / K;
    const int lane = lane_id();
    const int grp = lane / K, pos = lane % K;
    float *acc_g = acc + grp * kAccDim;

    int base = b + batch0 * 32;
    int nxt_c = 0;
    float nxt_w = 0.f;
    if (base + lane < e) {
        nxt_c = ld_stream_i32(idx + base + lane);
        nxt_w = ld_stream_f32(val + base + lane);
    }
    for (; base < e; base += stride * 32) {
        const int n = min(32, e - base);
        const int my_c = nxt_c;
        const float my_w = nxt_w;
        const int nb = base + stride * 32;  // prefetch the next batch of CSR entries
        nxt_c = 0;
        nxt_w = 0.f;
        if (nb + lane < e) {
            nxt_c = ld_stream_i32(idx + nb + lane);
            nxt_w = ld_stream_f32(val + nb + lane);
        }
        for (int j = 0; j < n; j += EPI * UNROLL) {
            float v[UNROLL], w[UNROLL];
            int s[UNROLL];
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                const int ej = j + u * EPI + grp;
                const int c = __shfl_sync(kFull, my_c, ej & 31);
                w[u] = __shfl_sync(kFull, my_w, ej & 31);
                s[u] = -1;
                v[u] = 0.f;
                if (ej < n) {
                    const size_t off = (size_t)c * K + pos;
                    v[u] = __ldg(cval + off);
                    s[u] = __ldg(csel + off);
                }
            }
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                // columns are distinct within a CBSR row and every edge slot has its own
                // accumulator copy, so this read-modify-write has no intra-instruction conflict
                if (s[u] >= 0) acc_g[s[u]] = fmaf(w[u], v[u], acc_g[s[u]]);
                __syncwarp();
            }
        }
    }
}

// k == 64: two entries per lane (8-byte value load, 2-byte selector load), one edge per instruction.
template <int UNROLL>
__device__ __forceinline__ void accumulate_k64(const int *__restrict__ idx, const float *__restrict__ val,
                                               const float *__restrict__ cval, const uint8_t *__restrict__ csel,
                                               float *acc, int b, int e, int batch0, int stride)
{
    const int lane = lane_id();
    int base = b + batch0 * 32;
    int nxt_c = 0;
    float nxt_w = 0.f;
    if (base + lane < e) {
        nxt_c = ld_stream_i32(idx + base + lane);
        nxt_w = ld_stream_f32(val + base + lane);
    }
    for (; base < e; base += stride * 32) {
        const int n = min(32, e - base);
        const int my_c = nxt_c;
        const float my_w = nxt_w;
        const int nb = base + stride * 32;
        nxt_c = 0;
        nxt_w = 0.f;
        if (nb + lane < e) {
            nxt_c = ld_stream_i32(idx + nb + lane);
            nxt_w = ld_stream_f32(val + nb + lane);
        }
        for (int j = 0; j < n; j += UNROLL) {
            float2 v[UNROLL];
            float w[UNROLL];
            int s[UNROLL];
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                const int ej = j + u;
                const int c = __shfl_sync(kFull, my_c, ej & 31);
                w[u] = __shfl_sync(kFull, my_w, ej & 31);
                s[u] = -1;
                v[u] = make_float2(0.f, 0.f);
                if (ej < n) {
                    const size_t off = (size_t)c * 64 + 2 * lane;
                    v[u] = __ldg(reinterpret_cast<const float2 *>(cval + off));
                    s[u] = __ldg(reinterpret_cast<const unsigned short *>(csel + off));
                }
            }
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                if (s[u] >= 0) {
                    const int s0 = s[u] & 0xff, s1 = s[u] >> 8;
                    acc[s0] = fmaf(w[u], v[u].x, acc[s0]);
                    acc[s1] = fmaf(w[u], v[u].y, acc[s1]);
                }
                __syncwarp();
            }
        }
    }
}

// any k in [1, 256]: one edge per step, lanes stride over the k entries.
__device__ __forceinline__ void accumulate_any_k(const int *__restrict__ idx, const float *__restrict__ val,
                                                 const float *__restrict__ cval, const uint8_t *__restrict__ csel,
                                                 float *acc, int k, int b, int e, int batch0, int stride)
{
    const int lane = lane_id();
    for (int base = b + batch0 * 32; base < e; base += stride * 32) {
        const int n = min(32, e - base);
        int my_c = 0;
        float my_w = 0.f;
        if (lane < n) {
            my_c = ld_stream_i32(idx + base + lane);
            my_w = ld_stream_f32(val + base + lane);
        }
        for (int j = 0; j < n; ++j) {
            const int c = __shfl_sync(kFull, my_c, j);
            const float w = __shfl_sync(kFull, my_w, j);
            for (int l = lane; l < k; l += 32) {
                const size_t off = (size_t)c * k + l;
                const int s = __ldg(csel + off);
                acc[s] = fmaf(w, __ldg(cval + off), acc[s]);
            }
            __syncwarp();
        }
    }
}

template <int K>
struct FwdTraits {
    static constexpr int kCopies = (K >= 8 && K <= 32) ? 32 / K : 1;  // accumulator copies per warp
};

template <int K>
__device__ __forceinline__ void accumulate_row(const int *idx, const float *val, const float *cval,
                                               const uint8_t *csel, float *acc, int k, int b, int e, int batch0,
                                               int stride)
{
    if constexpr (K == 32) accumulate_small_k<32, 8>(idx, val, cval, csel, acc, b, e, batch0, stride);
    else if constexpr (K == 16) accumulate_small_k<16, 8>(idx, val, cval, csel, acc, b, e, batch0, stride);
    else if constexpr (K == 8) accumulate_small_k<8, 8>(idx, val, cval, csel, acc, b, e, batch0, stride);
    else if constexpr (K == 64) accumulate_k64<4>(idx, val, cval, csel, acc, b, e, batch0, stride);
    else accumulate_any_k(idx, val, cval, csel, acc, k, b, e, batch0, stride);
}

// Row epilogue of the warp-owned path: sum the accumulator copies, re-zero them, apply the
// fused divisor and write the row once.
template <int COPIES>
__device__ __forceinline__ void write_row(float *acc, float *__restrict__ out_row, int dim, bool has_div, float div)
{
    const int lane = lane_id();
    float4 *acc4 = reinterpret_cast<float4 *>(acc);
    if (dim == kAccDim) {
        float4 a0 = acc4[lane], a1 = acc4[32 + lane];
        const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
        acc4[lane] = z;
        acc4[32 + lane] = z;
#pragma unroll
        for (int g = 1; g < COPIES; ++g) {
            const float4 b0 = acc4[g * 64 + lane], b1 = acc4[g * 64 + 32 + lane];
            a0.x += b0.x; a0.y += b0.y; a0.z += b0.z; a0.w += b0.w;
            a1.x += b1.x; a1.y += b1.y; a1.z += b1.z; a1.w += b1.w;
            acc4[g * 64 + lane] = z;
            acc4[g * 64 + 32 + lane] = z;
        }
        if (has_div) {
            a0.x /= div; a0.y /= div; a0.z /= div; a0.w /= div;
            a1.x /= div; a1.y /= div; a1.z /= div; a1.w /= div;
        }
        st_stream_f32x4(out_row + 4 * lane, a0);
        st_stream_f32x4(out_row + 128 + 4 * lane, a1);
    } else {
        for (int j = lane; j < kAccDim; j += 32) {
            float a = acc[j];
            acc[j] = 0.f;
#pragma unroll
            for (int g = 1; g < COPIES; ++g) {
                a += acc[g * kAccDim + j];
                acc[g * kAccDim + j] = 0.f;
            }
            if (has_div) a /= div;
            if (j < dim) out_row[j] = a;
        }
    }
    __syncwarp();
}

// ---------------------------------------------------------------------------------
// Main kernel: persistent grid, one warp per output row, dynamic row scheduling.
// ---------------------------------------------------------------------------------
template <int K>
__global__ void __launch_bounds__(kFwdThreads, 4)
spgemm_fwd_kernel(const int *__restrict__ row_begin, const int *__restrict__ row_end,
                  const int *__restrict__ idx, const float *__restrict__ val,
                  const float *__restrict__ cval, const uint8_t *__restrict__ csel, float *__restrict__ out,
                  int n_rows, int dim, int k, const float *__restrict__ row_div, SchedWorkspace *ws,
                  int *__restrict__ long_rows, int rows_per_grab)
{
    constexpr int COPIES = FwdTraits<K>::kCopies;
    extern __shared__ __align__(16) float smem[];
    const int lane = lane_id();
    float *acc = smem + (threadIdx.x >> 5) * (COPIES * kAccDim);
    for (int i = lane; i < COPIES * kAccDim; i += 32) acc[i] = 0.f;
    __syncwarp();

    for (;;) {
        int first = 0;
        if (lane == 0) first = atomicAdd(&ws->row_counter, rows_per_grab);
        first = __shfl_sync(kFull, first, 0);
        if (first >= n_rows) break;
        const int nr = min(rows_per_grab, n_rows - first);
        int rb = 0, re = 0;
        if (lane < nr) {
            rb = __ldg(row_begin + first + lane);
            re = __ldg(row_end + first + lane);
        }
        for (int i = 0; i < nr; ++i) {
            const int r = first + i;
            const int b = __shfl_sync(kFull, rb, i), e = __shfl_sync(kFull, re, i);
            if (e - b > kLongRow) {
                if (lane == 0) long_rows[atomicAdd(&ws->long_count, 1)] = r;
                continue;
            }
            if (e > b) accumulate_row<K>(idx, val, cval, csel, acc, k, b, e, 0, 1);
            const bool has_div = row_div != nullptr;
            write_row<COPIES>(acc, out + (size_t)r * dim, dim, has_div, has_div ? __ldg(row_div + r) : 1.f);
        }
    }
}

// ---------------------------------------------------------------------------------
// Long rows: one CTA per row, warps take alternating 32-edge batches, partial
// accumulators are summed across warps in a fixed order.
// ---------------------------------------------------------------------------------
template <int K>
__global__ void __launch_bounds__(kLongThreads, 1)
spgemm_fwd_long_kernel(const int *__restrict__ row_begin, const int *__restrict__ row_end,
                       const int *__restrict__ idx, const float *__restrict__ val,
                       const float *__restrict__ cval, const uint8_t *__restrict__ csel, float *__restrict__ out,
                       int dim, int k, const float *__restrict__ row_div, SchedWorkspace *ws,
                       const int *__restrict__ long_rows)
{
    constexpr int COPIES = FwdTraits<K>::kCopies;
    extern __shared__ __align__(16) float smem[];
    __shared__ int s_item;
    const int warp = threadIdx.x >> 5;
    float *acc = smem + warp * (COPIES * kAccDim);
    for (int i = threadIdx.x; i < kLongWarps * COPIES * kAccDim; i += kLongThreads) smem[i] = 0.f;
    const int n_long = ws->long_count;
    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) s_item = atomicAdd(&ws->long_counter, 1);
        __syncthreads();
        const int item = s_item;
        if (item >= n_long) break;
        const int r = long_rows[item];
        const int b = row_begin[r], e = row_end[r];
        accumulate_row<K>(idx, val, cval, csel, acc, k, b, e, warp, kLongWarps);
        __syncthreads();
        for (int j = threadIdx.x; j < kAccDim; j += kLongThreads) {
            float a = 0.f;
            for (int w = 0; w < kLongWarps * COPIES; ++w) {
                a += smem[w * kAccDim + j];
                smem[w * kAccDim + j] = 0.f;
            }
            if (row_div != nullptr) a /= row_div[r];
            if (j < dim) out[(size_t)r * dim + j] = a;
        }
    }
}

// ---------------------------------------------------------------------------------
// Host side
// ---------------------------------------------------------------------------------
struct DeviceInfo {
    int sms = 0;
};
static DeviceInfo device_info()
{
    DeviceInfo d;
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&d.sms, cudaDevAttrMultiProcessorCount, dev);
    if (d.sms <= 0) d.sms = kNumSMsB200;
    return d;
}

int pick_rows_per_grab(int64_t n_rows, int64_t n_edges, int total_warps)
{
    const int64_t avg = n_rows > 0 ? (n_edges + n_rows - 1) / n_rows : 1;
    int64_t by_balance = n_rows / (8 * (int64_t)total_warps);  // keep >= 8 grabs per warp
    int64_t by_work = 2048 / (avg > 0 ? avg : 1);              // <= ~2048 edges per grab
    int64_t g = by_balance < by_work ? by_balance : by_work;
    if (g < 1) g = 1;
    if (g > 32) g = 32;
    return (int)g;
}

template <int K>
static cudaError_t launch_fwd(const int *row_begin, const int *row_end, const int *idx, const float *val,
                              const float *cval, const uint8_t *csel, float *out, int64_t n_rows, int64_t n_edges,
                              int dim, int k, const float *row_div, SchedWorkspace *ws, cudaStream_t stream)
{
    constexpr int COPIES = FwdTraits<K>::kCopies;
    const size_t smem_main = (size_t)kFwdWarps * COPIES * kAccDim * sizeof(float);
    const size_t smem_long = (size_t)kLongWarps * COPIES * kAccDim * sizeof(float);
    static bool configured = false;  // per template instance
    static int blocks_per_sm = 1;
    if (!configured) {
        cudaFuncSetAttribute(spgemm_fwd_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_main);
        cudaFuncSetAttribute(spgemm_fwd_long_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_long);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, spgemm_fwd_kernel<K>, kFwdThreads, smem_main);
        if (blocks_per_sm < 1) blocks_per_sm = 1;
        configured = true;
    }
    const DeviceInfo di = device_info();
    int *long_rows = reinterpret_cast<int *>(ws + 1);
    cudaError_t err = cudaMemsetAsync(ws, 0, sizeof(SchedWorkspace), stream);
    if (err != cudaSuccess) return err;
    const int grid = di.sms * blocks_per_sm;
    const int rpg = pick_rows_per_grab(n_rows, n_edges, grid * kFwdWarps);
    spgemm_fwd_kernel<K><<<grid, kFwdThreads, smem_main, stream>>>(row_begin, row_end, idx, val, cval, csel, out,
                                                                   (int)n_rows, dim, k, row_div, ws, long_rows, rpg);
    err = cudaGetLastError();
    if (err != cudaSuccess) return err;
    spgemm_fwd_long_kernel<K><<<di.sms, kLongThreads, smem_long, stream>>>(row_begin, row_end, idx, val, cval, csel,
                                                                          out, dim, k, row_div, ws, long_rows);
    return cudaGetLastError();
}

}  // namespace maxk

using namespace maxk;

extern "C" size_t maxk_spgemm_workspace_bytes(int64_t n_rows)
{
    if (n_rows < 0) n_rows = 0;
    return sizeof(SchedWorkspace) + sizeof(int) * (size_t)n_rows + 16;
}

extern "C" int maxk_spgemm_forward(const int32_t *row_begin, const int32_t *row_end, const int32_t *indices,
                                   const float *values, const float *cbsr_val, const uint8_t *cbsr_sel, float *out,
                                   int64_t n_rows, int64_t n_edges, int dim, int k, const float *row_div,
                                   void *workspace, size_t workspace_bytes, maxk_stream_t stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    if (dim < 1 || dim > kAccDim) return MAXK_ERR_BAD_DIM;
    if (k < 1 || k > dim) return MAXK_ERR_BAD_K;
    if (n_rows < 0 || n_edges < 0 || n_rows > INT32_MAX || n_edges > INT32_MAX) return MAXK_ERR_SIZE;
    if (n_rows == 0) return MAXK_OK;
    if (!row_begin || !row_end || !out || !workspace) return MAXK_ERR_NULL;
    if (n_edges > 0 && (!indices || !values || !cbsr_val || !cbsr_sel)) return MAXK_ERR_NULL;
    if (workspace_bytes < maxk_spgemm_workspace_bytes(n_rows)) return MAXK_ERR_WORKSPACE;
    if (((uintptr_t)out | (uintptr_t)cbsr_val | (uintptr_t)workspace) & 15) return MAXK_ERR_ALIGN;
    SchedWorkspace *ws = reinterpret_cast<SchedWorkspace *>(workspace);
    cudaError_t err;
    switch (k) {
        case 8: err = launch_fwd<8>(row_begin, row_end, indices, values, cbsr_val, cbsr_sel, out, n_rows, n_edges, dim, k, row_div, ws, stream); break;
        case 16: err = launch_fwd<16>(row_begin, row_end, indices, values, cbsr_val, cbsr_sel, out, n_rows, n_edges, dim, k, row_div, ws, stream); break;
        case 32: err = launch_fwd<32>(row_begin, row_end, indices, values, cbsr_val, cbsr_sel, out, n_rows, n_edges, dim, k, row_div, ws, stream); break;
        case 64: err = launch_fwd<64>(row_begin, row_end, indices, values, cbsr_val, cbsr_sel, out, n_rows, n_edges, dim, k, row_div, ws, stream); break;
        default: err = launch_fwd<0>(row_begin, row_end, indices, values, cbsr_val, cbsr_sel, out, n_rows, n_edges, dim, k, row_div, ws, stream); break;
    }
    return status_from_cuda(err);
}
