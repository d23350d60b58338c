// spgemm_fwd.cu -- forward row-wise-product SpGEMM, CSR adjacency x CBSR features (sm_100a).
//
// Replaces spmm_kernel_opt2_sparse_v3 (reference kernels/spmm_maxk.cu:17-106).  Design
// (DESIGN.md "Forward SpGEMM"):
//   * one warp OWNS one output row: the whole edge list of the row is reduced by that warp
//     into private shared-memory accumulators and the row is written exactly once.  No global
//     atomics, no output memset, fixed reduction order (the reference flushes 256 global
//     atomics per <=64-edge segment, spmm_maxk.cu:101-105, into a torch::zeros output,
//     cuda_kernel_bindings.cpp:71).
//   * CSR indices/values are read as coalesced 128-byte streaming loads, 32 edges per warp
//     instruction, prefetched one batch ahead, and handed to the lanes by shuffles (the
//     reference issues one scalar broadcast __ldg per edge, spmm_maxk.cu:72-73).
//   * neighbour CBSR rows are gathered with 16-byte loads, EPI = 128/k whole rows per warp
//     instruction, UNROLL instructions in flight before the first accumulate; the accumulators
//     use the banked per-edge-slot layout of maxk_common.cuh (Lay<K>), which keeps the LSU
//     data pipe -- the measured limiter -- at the fewest wavefronts per edge.  All 32 lanes work
//     for every k (the reference idles 1 - k/32 of its warps for k < 32, spmm_maxk.cu:27,64).
//   * rows are handed out dynamically (one global counter), rows longer than kLongRow are
//     deferred to a whole-CTA kernel, so skewed graphs do not serialise on one warp.
//   * the degree normalisation the reference does in a separate PyTorch pass
//     (maxk_spgemm_function.py:86) is fused into the row epilogue.
#include "maxk_common.cuh"

namespace maxk {

constexpr int kFwdThreads = 256;
constexpr int kFwdWarps = kFwdThreads / 32;
constexpr int kLongThreads = 512;
constexpr int kLongWarps = kLongThreads / 32;

template <int EPL> struct EntryLoad;
template <> struct EntryLoad<4> {
    float v[4];
    uint32_t s;
    __device__ __forceinline__ void load(const float *cval, const uint8_t *csel, size_t off, uint64_t keep)
    {
        const float4 t = ld_keep_f32x4(cval + off, keep);
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
        s = ld_keep_u32(csel + off, keep);
    }
};
template <> struct EntryLoad<2> {
    float v[2];
    uint32_t s;
    __device__ __forceinline__ void load(const float *cval, const uint8_t *csel, size_t off, uint64_t keep)
    {
        const float2 t = ld_keep_f32x2(cval + off, keep);
        v[0] = t.x; v[1] = t.y;
        s = ld_keep_u16(csel + off, keep);
    }
};

// ---------------------------------------------------------------------------------
// Edge accumulation for one row segment [b, e), visiting 32-edge batches
// batch0, batch0+stride, ...   Fast path: k in {8, 16, 32, 64}.
// ---------------------------------------------------------------------------------
template <int K, int UNROLL, bool PREFETCHED>
__device__ __forceinline__ void accumulate_fast(const int *__restrict__ idx, const float *__restrict__ val,
                                                const float *__restrict__ cval, const uint8_t *__restrict__ csel,
                                                float *acc, int2 *cw, int b, int e, int batch0, int stride,
                                                int first_c, float first_w)
{
    using LY = Lay<K>;
    constexpr int EPL = LY::EPL, L = LY::L, EPI = LY::EPI;
    const int lane = lane_id();
    const int q = lane / L, t = lane % L;
    float *acc_q = acc + q * L;
    const uint64_t keep = policy_evict_last();     // CBSR rows are re-used ~degree times: keep them in L2

    int base = b + batch0 * 32;
    int nxt_c = first_c;       // PREFETCHED: the caller already loaded the first batch of this row
    float nxt_w = first_w;
    if (!PREFETCHED) {
        nxt_c = 0;
        nxt_w = 0.f;
        if (base + lane < e) {
            nxt_c = ld_stream_i32(idx + base + lane);
            nxt_w = ld_stream_f32(val + base + lane);
        }
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
        // k >= 32: the batch's (source id, edge value) pairs go through a 256-byte shared-memory row, one
        // 8-byte broadcast load per instruction instead of two shuffles (measured -2.5 %; for k < 32 the
        // extra synchronisation costs more than it saves, so those keep the shuffles)
        constexpr bool kSmemBroadcast = K >= 32;
        if (kSmemBroadcast) {
            __syncwarp();
            cw[lane] = make_int2(my_c, __float_as_int(my_w));
            __syncwarp();
        }
        for (int j = 0; j < n; j += EPI * UNROLL) {
            EntryLoad<EPL> ent[UNROLL];
            float w[UNROLL];
            bool ok[UNROLL];
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                const int ej = j + u * EPI + q;
                int2 cwj;
                if (kSmemBroadcast) {
                    cwj = cw[ej & 31];
                } else {
                    cwj.x = __shfl_sync(kFullMask, my_c, ej & 31);
                    cwj.y = __shfl_sync(kFullMask, __float_as_int(my_w), ej & 31);
                }
                w[u] = __int_as_float(cwj.y);
                ok[u] = ej < n;
                if (ok[u]) ent[u].load(cval, csel, (size_t)cwj.x * K + EPL * t, keep);
            }
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                if (ok[u]) {
                    // the EPL columns of a lane and the columns of the other lanes of this edge
                    // slot are all distinct (one CBSR row), other slots use other banks: the
                    // EPL read-modify-writes are independent, so read all, then write all.
                    int o[EPL];
                    float a[EPL];
#pragma unroll
                    for (int i = 0; i < EPL; ++i) {
                        o[i] = LY::word((ent[u].s >> (8 * i)) & 0xff);
                        a[i] = acc_q[o[i]];
                    }
#pragma unroll
                    for (int i = 0; i < EPL; ++i) acc_q[o[i]] = fmaf(w[u], ent[u].v[i], a[i]);
                }
                __syncwarp();  // next step may touch the same column of the same copy from another lane
            }
        }
    }
}

// any k in [1, 256]: one edge per step, lanes stride over the k entries (natural layout, one copy).
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
            const int c = __shfl_sync(kFullMask, my_c, j);
            const float w = __shfl_sync(kFullMask, my_w, j);
            for (int l = lane; l < k; l += 32) {
                const size_t off = (size_t)c * k + l;
                const int s = __ldg(csel + off);
                acc[s] = fmaf(w, __ldg(cval + off), acc[s]);
            }
            __syncwarp();
        }
    }
}

template <int K, bool PREFETCHED>
__device__ __forceinline__ void accumulate_row(const int *idx, const float *val, const float *cval,
                                               const uint8_t *csel, float *acc, int2 *cw, int k, int b, int e,
                                               int batch0, int stride, int first_c, float first_w)
{
    if constexpr (Lay<K>::kFast)
        accumulate_fast<K, 4, PREFETCHED>(idx, val, cval, csel, acc, cw, b, e, batch0, stride, first_c, first_w);
    else accumulate_any_k(idx, val, cval, csel, acc, k, b, e, batch0, stride);
}

// Column `col` summed over the EPI copies of `n_sets` accumulator sets (sets are kWords apart),
// read with a lane-skewed copy order so that the 32 lanes hit 32 different banks.
template <int K>
__device__ __forceinline__ float sum_copies(const float *acc, int col, int lane, int n_sets)
{
    using LY = Lay<K>;
    const int w0 = LY::word(col);
    float a = 0.f;
    for (int s = 0; s < n_sets; ++s) {
#pragma unroll
        for (int q = 0; q < LY::EPI; ++q) {
            const int qq = (q + lane / LY::L) % LY::EPI;
            a += acc[s * LY::kWords + w0 + qq * LY::L];
        }
    }
    return a;
}

// Row epilogue of the warp-owned path: sum the copies, re-zero them, fused divisor, one write.
// Lane l owns columns l + 32 n.  word(l + 32 n) = word(l) + n * (32 / L) * 32, so every address is
// a per-lane base + a compile-time offset: the 8 * EPI loads need no address arithmetic.
template <int K>
__device__ __forceinline__ void write_row(float *acc, float *__restrict__ out_row, int dim, bool has_div, float div)
{
    using LY = Lay<K>;
    const int lane = lane_id();
    constexpr int kColStep = (32 / LY::L) * 32;      // words between column c and column c + 32
    const float *base = acc + LY::word(lane);
    float o[kAccDim / 32];
#pragma unroll
    for (int n = 0; n < kAccDim / 32; ++n) o[n] = 0.f;
#pragma unroll
    for (int q = 0; q < LY::EPI; ++q) {
        const float *bq = base + ((q + lane / LY::L) % LY::EPI) * LY::L;   // lane-skewed copy: 32 distinct banks
#pragma unroll
        for (int n = 0; n < kAccDim / 32; ++n) o[n] += bq[n * kColStep];
    }
    __syncwarp();
    float4 *acc4 = reinterpret_cast<float4 *>(acc);
#pragma unroll
    for (int i = 0; i < LY::kWords / 128; ++i) acc4[lane + 32 * i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (has_div) {                                    // warp-uniform
        const float r = 1.0f / div;
        if (recip_usable(r)) {                        // normal divisor (degrees are >= 1): reciprocal path
#pragma unroll
            for (int n = 0; n < kAccDim / 32; ++n) o[n] = div_by_recip(o[n], div, r);
        } else {                                      // zero / infinite / NaN divisor: plain IEEE division
#pragma unroll
            for (int n = 0; n < kAccDim / 32; ++n) o[n] = o[n] / div;
        }
    }
    if (dim == kAccDim) {
#pragma unroll
        for (int n = 0; n < kAccDim / 32; ++n) st_stream_f32(out_row + lane + 32 * n, o[n]);
    } else {
#pragma unroll
        for (int n = 0; n < kAccDim / 32; ++n)
            if (lane + 32 * n < dim) st_stream_f32(out_row + lane + 32 * n, o[n]);
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
    using LY = Lay<K>;
    extern __shared__ __align__(16) float smem[];
    __shared__ int2 s_cw[kFwdWarps][32];
    const int lane = lane_id();
    float *acc = smem + (threadIdx.x >> 5) * LY::kWords;
    int2 *cw = s_cw[threadIdx.x >> 5];
    for (int i = lane; i < LY::kWords; i += 32) acc[i] = 0.f;
    __syncwarp();

    for (;;) {
        int first = 0;
        if (lane == 0) first = atomicAdd(&ws->row_counter, rows_per_grab);
        first = __shfl_sync(kFullMask, first, 0);
        if (first >= n_rows) break;
        const int nr = min(rows_per_grab, n_rows - first);
        int rb = 0, re = 0;
        if (lane < nr) {
            rb = __ldg(row_begin + first + lane);
            re = __ldg(row_end + first + lane);
        }
        // software pipeline over the rows of the grab: the first CSR batch of row i+1 is in flight
        // while row i is reduced (low-degree graphs are otherwise one DRAM latency per row)
        int nb = __shfl_sync(kFullMask, rb, 0), ne = __shfl_sync(kFullMask, re, 0);
        int pc = 0;
        float pw = 0.f;
        if (Lay<K>::kFast && ne - nb <= kLongRow && nb + lane < ne) {
            pc = ld_stream_i32(idx + nb + lane);
            pw = ld_stream_f32(val + nb + lane);
        }
        for (int i = 0; i < nr; ++i) {
            const int r = first + i;
            const int b = nb, e = ne;
            const int cur_c = pc;
            const float cur_w = pw;
            if (i + 1 < nr) {
                nb = __shfl_sync(kFullMask, rb, i + 1);
                ne = __shfl_sync(kFullMask, re, i + 1);
                pc = 0;
                pw = 0.f;
                if (Lay<K>::kFast && ne - nb <= kLongRow && nb + lane < ne) {
                    pc = ld_stream_i32(idx + nb + lane);
                    pw = ld_stream_f32(val + nb + lane);
                }
            }
            if (e - b > kLongRow) {
                if (lane == 0) long_rows[atomicAdd(&ws->long_count, 1)] = r;
                continue;
            }
            if (e > b) accumulate_row<K, true>(idx, val, cval, csel, acc, cw, k, b, e, 0, 1, cur_c, cur_w);
            const bool has_div = row_div != nullptr;
            write_row<K>(acc, out + (size_t)r * dim, dim, has_div, has_div ? __ldg(row_div + r) : 1.f);
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
    using LY = Lay<K>;
    extern __shared__ __align__(16) float smem[];
    __shared__ int s_item;
    __shared__ int2 s_cw[kLongWarps][32];
    const int warp = threadIdx.x >> 5;
    float *acc = smem + warp * LY::kWords;
    int2 *cw = s_cw[warp];
    for (int i = threadIdx.x; i < kLongWarps * LY::kWords; i += kLongThreads) smem[i] = 0.f;
    const int n_long = ws->long_count;
    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) s_item = atomicAdd(&ws->long_counter, 1);
        __syncthreads();
        const int item = s_item;
        if (item >= n_long) break;
        const int r = long_rows[item];
        const int b = row_begin[r], e = row_end[r];
        accumulate_row<K, false>(idx, val, cval, csel, acc, cw, k, b, e, warp, kLongWarps, 0, 0.f);
        __syncthreads();
        float o = 0.f;
        if (threadIdx.x < kAccDim) o = sum_copies<K>(smem, threadIdx.x, threadIdx.x & 31, kLongWarps);
        __syncthreads();
        for (int i = threadIdx.x; i < kLongWarps * LY::kWords; i += kLongThreads) smem[i] = 0.f;
        if (threadIdx.x < dim) {
            if (row_div != nullptr) o = div_guarded(o, row_div[r]);
            out[(size_t)r * dim + threadIdx.x] = o;
        }
    }
}

// ---------------------------------------------------------------------------------
// Host side
// ---------------------------------------------------------------------------------
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
    using LY = Lay<K>;
    const size_t smem_main = (size_t)kFwdWarps * LY::kWords * sizeof(float);
    const size_t smem_long = (size_t)kLongWarps * LY::kWords * sizeof(float);
    static LaunchConfig cache[kMaxCachedDevices];  // per template instance and device
    int dev = 0;
    cudaError_t err = cudaGetDevice(&dev);
    if (err != cudaSuccess) return err;
    LaunchConfig uncached = {false, 1, kNumSMsB200};
    LaunchConfig &cfg = (dev >= 0 && dev < kMaxCachedDevices) ? cache[dev] : uncached;
    if (!cfg.configured) {
        err = configure_pair(cfg, spgemm_fwd_kernel<K>, spgemm_fwd_long_kernel<K>, kFwdThreads, smem_main, smem_long);
        if (err != cudaSuccess) return err;
    }
    const int sms = cfg.sms;
    int *long_rows = reinterpret_cast<int *>(ws + 1);
    err = cudaMemsetAsync(ws, 0, sizeof(SchedWorkspace), stream);
    if (err != cudaSuccess) return err;
    const int grid = sms * cfg.blocks_per_sm;
    const int rpg = pick_rows_per_grab(n_rows, n_edges, grid * kFwdWarps);
    spgemm_fwd_kernel<K><<<grid, kFwdThreads, smem_main, stream>>>(row_begin, row_end, idx, val, cval, csel, out,
                                                                   (int)n_rows, dim, k, row_div, ws, long_rows, rpg);
    err = cudaGetLastError();
    if (err != cudaSuccess) return err;
    spgemm_fwd_long_kernel<K><<<sms, kLongThreads, smem_long, stream>>>(row_begin, row_end, idx, val, cval, csel, out,
                                                                       dim, k, row_div, ws, long_rows);
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
    if (k < 1 || k > kAccDim) return MAXK_ERR_BAD_K;
    if (n_rows < 0 || n_edges < 0 || n_rows > INT32_MAX || n_edges > INT32_MAX) return MAXK_ERR_SIZE;
    if (n_rows == 0) return MAXK_OK;
    if (!row_begin || !row_end || !out || !workspace) return MAXK_ERR_NULL;
    if (n_edges > 0 && (!indices || !values || !cbsr_val || !cbsr_sel)) return MAXK_ERR_NULL;
    if (workspace_bytes < maxk_spgemm_workspace_bytes(n_rows)) return MAXK_ERR_WORKSPACE;
    if (((uintptr_t)cbsr_val | (uintptr_t)workspace) & 15) return MAXK_ERR_ALIGN;
    if (((uintptr_t)cbsr_sel | (uintptr_t)out) & 3) return MAXK_ERR_ALIGN;
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
