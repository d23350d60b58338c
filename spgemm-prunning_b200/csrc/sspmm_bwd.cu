// sspmm_bwd.cu -- backward outer-product SSpMM (sm_100a).
//
// Replaces spmm_kernel_opt2_sparse_backward_v3 (reference kernels/spmm_maxk_backward.cu:15-115):
//   gs[c, l] += val[e] * g[r, sel[c,l]]    for every edge e=(r,c), l<k
// Design (DESIGN.md "Backward SSpMM"):
//   * outer-product form kept: a warp stages its source row g[r,0:256] (1 KiB, coalesced,
//     fused /row_div) in shared memory once and walks the row's edges; the gather form would
//     re-read up to 1 KiB of g per EDGE.
//   * the E*k scalar fp32 atomics of the reference (spmm_maxk_backward.cu:80,101) become
//     E*k/4 16-byte vector reductions (red.global.add.v4.f32, SASS REDG.E.ADD.F32x4):
//     a lane owns 4 consecutive entries of one destination row, k/4 lanes cover an edge and
//     128/k edges are processed per warp instruction.
//   * selectors are fetched as one 4-byte load per lane (the reference does one byte load
//     per lane per edge), CSR indices/values as coalesced 128-byte streaming loads.
//   * output zero-fill is part of the call (cudaMemsetAsync on the stream), rows are
//     scheduled dynamically, long rows are spread over a whole CTA.
#include "maxk_common.cuh"

namespace maxk {

constexpr int kBwdThreads = 256;
constexpr int kBwdWarps = kBwdThreads / 32;
constexpr int kBwdLongThreads = 512;
constexpr int kBwdLongWarps = kBwdLongThreads / 32;
constexpr unsigned kFullB = 0xffffffffu;

// Stage one row of g (dim <= 256, fused division) into the warp's shared-memory slot in the
// banked layout of Lay<K>: every edge slot q gets its own copy living in banks [q*L, q*L+L), so
// lanes of different edge slots never collide when they look their columns up.  The copies are
// written in a lane-skewed order (32 distinct banks per store instruction).
__device__ __forceinline__ void load_g_row(const float *__restrict__ g_row, int dim, float (&gv)[kAccDim / 32])
{
    const int lane = lane_id();
#pragma unroll
    for (int n = 0; n < kAccDim / 32; ++n) {
        const int col = lane + 32 * n;
        gv[n] = col < dim ? ld_stream_f32(g_row + col) : 0.f;
    }
}

template <int K>
__device__ __forceinline__ void stage_row(const float (&gv)[kAccDim / 32], float *gsm, bool has_div, float div)
{
    using LY = Lay<K>;
    const int lane = lane_id();
    constexpr int kColStep = (32 / LY::L) * 32;      // words between column c and column c + 32
    float a[kAccDim / 32];
#pragma unroll
    for (int n = 0; n < kAccDim / 32; ++n) a[n] = gv[n];
    if (has_div) {                                    // warp-uniform; one reciprocal per row (maxk_common.cuh)
        const float r = 1.0f / div;
        if (recip_usable(r)) {                        // normal divisor (degrees are >= 1): reciprocal path
#pragma unroll
            for (int n = 0; n < kAccDim / 32; ++n) a[n] = div_by_recip(a[n], div, r);
        } else {                                      // zero / infinite / NaN divisor: plain IEEE division
#pragma unroll
            for (int n = 0; n < kAccDim / 32; ++n) a[n] = a[n] / div;
        }
    }
    float *base = gsm + LY::word(lane);
#pragma unroll
    for (int q = 0; q < LY::EPI; ++q) {
        float *bq = base + ((q + lane / LY::L) % LY::EPI) * LY::L;         // lane-skewed copy: 32 distinct banks
#pragma unroll
        for (int n = 0; n < kAccDim / 32; ++n) bq[n * kColStep] = a[n];
    }
    __syncwarp();
}

template <int EPL> struct SelLoad;
template <> struct SelLoad<4> {
    __device__ static __forceinline__ uint32_t load(const uint8_t *p, uint64_t keep) { return ld_keep_u32(p, keep); }
    __device__ static __forceinline__ void red(float *p, const float *x, uint64_t keep) { red_add_f32x4(p, x[0], x[1], x[2], x[3], keep); }
};
template <> struct SelLoad<2> {
    __device__ static __forceinline__ uint32_t load(const uint8_t *p, uint64_t keep) { return ld_keep_u16(p, keep); }
    __device__ static __forceinline__ void red(float *p, const float *x, uint64_t keep) { red_add_f32x2(p, x[0], x[1], keep); }
};

// Fast path, k in {8, 16, 32, 64}: a lane owns EPL consecutive entries of one destination row,
// L = k/EPL lanes cover an edge, EPI = 32/L edges per warp instruction; one vector reduction
// (16 B, or 8 B for k = 8) per lane per edge.
// KS = entries per CBSR row.  KS == K for k in {8, 16, 32, 64}; k = 96 / 128 run the k = 32 lane layout over
// KS / 32 chunks of every row (each chunk is one full 128-byte line of gs, one vector reduction per lane).
template <int K, int KS, int UNROLL, bool PREFETCHED>
__device__ __forceinline__ void scatter_fast(const int *__restrict__ idx, const float *__restrict__ val,
                                             const uint8_t *__restrict__ csel, float *__restrict__ gs,
                                             const float *gsm, int b, int e, int batch0, int stride, int first_c,
                                             float first_w)
{
    using LY = Lay<K>;
    constexpr int EPL = LY::EPL, L = LY::L, EPI = LY::EPI;
    const int lane = lane_id();
    const int q = lane / L, t = lane % L;
    const float *gsm_q = gsm + q * L;
    const uint64_t keep = policy_evict_last();     // selectors and the sampled gradient are re-used: keep them in L2

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
        const int nb = base + stride * 32;
        nxt_c = 0;
        nxt_w = 0.f;
        if (nb + lane < e) {
            nxt_c = ld_stream_i32(idx + nb + lane);
            nxt_w = ld_stream_f32(val + nb + lane);
        }
        for (int ch = 0; ch < KS / K; ++ch)
        for (int j = 0; j < n; j += EPI * UNROLL) {
            uint32_t s[UNROLL];
            float w[UNROLL];
            size_t off[UNROLL];
            bool ok[UNROLL];
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                const int ej = j + u * EPI + q;
                const int c = __shfl_sync(kFullB, my_c, ej & 31);
                w[u] = __shfl_sync(kFullB, my_w, ej & 31);
                ok[u] = ej < n;
                off[u] = (size_t)c * KS + ch * K + EPL * t;
                s[u] = 0;
                if (ok[u]) s[u] = SelLoad<EPL>::load(csel + off[u], keep);
            }
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                if (ok[u]) {
                    float x[EPL];
#pragma unroll
                    for (int i = 0; i < EPL; ++i) x[i] = w[u] * gsm_q[LY::word((s[u] >> (8 * i)) & 0xff)];
                    SelLoad<EPL>::red(gs + off[u], x, keep);
                }
            }
        }
    }
}

// any k: scalar reductions, one edge per step.
__device__ __forceinline__ void scatter_any_k(const int *__restrict__ idx, const float *__restrict__ val,
                                              const uint8_t *__restrict__ csel, float *__restrict__ gs,
                                              const float *gsm, int k, int b, int e, int batch0, int stride)
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
            const int c = __shfl_sync(kFullB, my_c, j);
            const float w = __shfl_sync(kFullB, my_w, j);
            for (int l = lane; l < k; l += 32) {
                const size_t off = (size_t)c * k + l;
                atomicAdd(gs + off, w * gsm[__ldg(csel + off)]);
            }
        }
    }
}

// lane layout used for a row width: k = 96 / 128 borrow the k = 32 layout (3 / 4 chunks per row)
template <int K> struct BwdLay {
    static constexpr int LK = (K == 96 || K == 128) ? 32 : K;
    using LY = Lay<LK>;
};

template <int K, bool PREFETCHED>
__device__ __forceinline__ void scatter_row(const int *idx, const float *val, const uint8_t *csel, float *gs,
                                            const float *gsm, int k, int b, int e, int batch0, int stride,
                                            int first_c, float first_w)
{
    if constexpr (BwdLay<K>::LY::kFast)
        scatter_fast<BwdLay<K>::LK, K, 4, PREFETCHED>(idx, val, csel, gs, gsm, b, e, batch0, stride, first_c, first_w);
    else scatter_any_k(idx, val, csel, gs, gsm, k, b, e, batch0, stride);
}

template <int K>
__global__ void __launch_bounds__(kBwdThreads, 4)
sspmm_bwd_kernel(const int *__restrict__ row_begin, const int *__restrict__ row_end,
                 const int *__restrict__ idx, const float *__restrict__ val, const float *__restrict__ g,
                 const uint8_t *__restrict__ csel, float *__restrict__ gs, int n_rows, int dim, int ld_g, int k,
                 const float *__restrict__ row_div, SchedWorkspace *ws, int *__restrict__ long_rows,
                 int rows_per_grab)
{
    extern __shared__ __align__(16) float smem[];
    const int lane = lane_id();
    float *gsm = smem + (threadIdx.x >> 5) * BwdLay<K>::LY::kWords;
    for (;;) {
        int first = 0;
        if (lane == 0) first = atomicAdd(&ws->row_counter, rows_per_grab);
        first = __shfl_sync(kFullB, first, 0);
        if (first >= n_rows) break;
        const int nr = min(rows_per_grab, n_rows - first);
        int rb = 0, re = 0;
        if (lane < nr) {
            rb = __ldg(row_begin + first + lane);
            re = __ldg(row_end + first + lane);
        }
        // software pipeline over the rows of the grab: the gradient row and the first CSR batch of
        // row i+1 are in flight while row i is scattered
        int nb = __shfl_sync(kFullB, rb, 0), ne = __shfl_sync(kFullB, re, 0);
        int pc = 0;
        float pw = 0.f;
        float pg[kAccDim / 32];
        if (ne > nb && ne - nb <= kLongRow) {
            load_g_row(g + (size_t)first * ld_g, dim, pg);
            if (BwdLay<K>::LY::kFast && nb + lane < ne) {
                pc = ld_stream_i32(idx + nb + lane);
                pw = ld_stream_f32(val + nb + lane);
            }
        }
        for (int i = 0; i < nr; ++i) {
            const int r = first + i;
            const int b = nb, e = ne;
            const int cur_c = pc;
            const float cur_w = pw;
            float cg[kAccDim / 32];
#pragma unroll
            for (int n = 0; n < kAccDim / 32; ++n) cg[n] = pg[n];
            if (i + 1 < nr) {
                nb = __shfl_sync(kFullB, rb, i + 1);
                ne = __shfl_sync(kFullB, re, i + 1);
                pc = 0;
                pw = 0.f;
                if (ne > nb && ne - nb <= kLongRow) {
                    load_g_row(g + (size_t)(r + 1) * ld_g, dim, pg);
                    if (BwdLay<K>::LY::kFast && nb + lane < ne) {
                        pc = ld_stream_i32(idx + nb + lane);
                        pw = ld_stream_f32(val + nb + lane);
                    }
                }
            }
            if (e <= b) continue;
            if (e - b > kLongRow) {
                if (lane == 0) long_rows[atomicAdd(&ws->long_count, 1)] = r;
                continue;
            }
            const bool has_div = row_div != nullptr;
            __syncwarp();
            stage_row<BwdLay<K>::LK>(cg, gsm, has_div, has_div ? __ldg(row_div + r) : 1.f);
            scatter_row<K, true>(idx, val, csel, gs, gsm, k, b, e, 0, 1, cur_c, cur_w);
        }
    }
}

template <int K>
__global__ void __launch_bounds__(kBwdLongThreads, 1)
sspmm_bwd_long_kernel(const int *__restrict__ row_begin, const int *__restrict__ row_end,
                      const int *__restrict__ idx, const float *__restrict__ val, const float *__restrict__ g,
                      const uint8_t *__restrict__ csel, float *__restrict__ gs, int dim, int ld_g, int k,
                      const float *__restrict__ row_div, SchedWorkspace *ws, const int *__restrict__ long_rows)
{
    extern __shared__ __align__(16) float smem[];
    __shared__ int s_item;
    const int warp = threadIdx.x >> 5;
    float *gsm = smem + warp * BwdLay<K>::LY::kWords;
    const int n_long = ws->long_count;
    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) s_item = atomicAdd(&ws->long_counter, 1);
        __syncthreads();
        const int item = s_item;
        if (item >= n_long) break;
        const int r = long_rows[item];
        const int b = row_begin[r], e = row_end[r];
        const bool has_div = row_div != nullptr;
        float gv[kAccDim / 32];
        load_g_row(g + (size_t)r * ld_g, dim, gv);
        stage_row<BwdLay<K>::LK>(gv, gsm, has_div, has_div ? __ldg(row_div + r) : 1.f);
        scatter_row<K, false>(idx, val, csel, gs, gsm, k, b, e, warp, kBwdLongWarps, 0, 0.f);
    }
}

static int pick_rows_per_grab(int64_t n_rows, int64_t n_edges, int total_warps)
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
static cudaError_t launch_bwd(const int *row_begin, const int *row_end, const int *idx, const float *val,
                              const float *g, const uint8_t *csel, float *gs, int64_t n_rows, int64_t n_dst,
                              int64_t n_edges, int dim, int ld_g, int k, const float *row_div, SchedWorkspace *ws,
                              bool zero_fill, cudaStream_t stream)
{
    const size_t smem_main = (size_t)kBwdWarps * BwdLay<K>::LY::kWords * sizeof(float);
    const size_t smem_long = (size_t)kBwdLongWarps * BwdLay<K>::LY::kWords * sizeof(float);
    static LaunchConfig cache[kMaxCachedDevices];  // per template instance and device
    int dev = 0;
    cudaError_t err = cudaGetDevice(&dev);
    if (err != cudaSuccess) return err;
    LaunchConfig uncached = {false, 1, kNumSMsB200};
    LaunchConfig &cfg = (dev >= 0 && dev < kMaxCachedDevices) ? cache[dev] : uncached;
    if (!cfg.configured) {
        err = configure_pair(cfg, sspmm_bwd_kernel<K>, sspmm_bwd_long_kernel<K>, kBwdThreads, smem_main, smem_long);
        if (err != cudaSuccess) return err;
    }
    const int sms = cfg.sms;
    int *long_rows = reinterpret_cast<int *>(ws + 1);
    err = cudaMemsetAsync(ws, 0, sizeof(SchedWorkspace), stream);
    if (err != cudaSuccess) return err;
    if (zero_fill) {
        err = cudaMemsetAsync(gs, 0, sizeof(float) * (size_t)n_dst * k, stream);
        if (err != cudaSuccess) return err;
    }
    const int grid = sms * cfg.blocks_per_sm;
    const int rpg = pick_rows_per_grab(n_rows, n_edges, grid * kBwdWarps);
    sspmm_bwd_kernel<K><<<grid, kBwdThreads, smem_main, stream>>>(row_begin, row_end, idx, val, g, csel, gs, (int)n_rows, dim,
                                                          ld_g, k, row_div, ws, long_rows, rpg);
    err = cudaGetLastError();
    if (err != cudaSuccess) return err;
    sspmm_bwd_long_kernel<K><<<sms, kBwdLongThreads, smem_long, stream>>>(row_begin, row_end, idx, val, g, csel, gs, dim, ld_g, k,
                                                                  row_div, ws, long_rows);
    return cudaGetLastError();
}

}  // namespace maxk

using namespace maxk;

// g rows may be strided (ld_g floats apart): the wide-feature path (wide.cu) reads 256-column blocks of a
// [n_rows, D > 256] gradient in place.
int maxk_backward_strided(bool zero_fill, const int32_t *row_begin, const int32_t *row_end, const int32_t *indices,
                                   const float *values, const float *g, int64_t ld_g, const uint8_t *cbsr_sel, float *gs,
                                   int64_t n_rows, int64_t n_dst, int64_t n_edges, int dim, int k,
                                   const float *row_div, void *workspace, size_t workspace_bytes,
                                   maxk_stream_t stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    if (dim < 1 || dim > kAccDim) return MAXK_ERR_BAD_DIM;
    if (k < 1 || k > kAccDim) return MAXK_ERR_BAD_K;
    if (n_rows < 0 || n_dst < 0 || n_edges < 0 || n_rows > INT32_MAX || n_edges > INT32_MAX || ld_g < dim || ld_g > INT32_MAX)
        return MAXK_ERR_SIZE;
    if (n_dst == 0) return MAXK_OK;
    if (!gs) return MAXK_ERR_NULL;
    if (n_rows == 0 || n_edges == 0)
        return zero_fill ? status_from_cuda(cudaMemsetAsync(gs, 0, sizeof(float) * (size_t)n_dst * k, stream)) : MAXK_OK;
    if (!row_begin || !row_end || !indices || !values || !g || !cbsr_sel || !workspace) return MAXK_ERR_NULL;
    if (workspace_bytes < maxk_spgemm_workspace_bytes(n_rows)) return MAXK_ERR_WORKSPACE;
    if (((uintptr_t)gs | (uintptr_t)workspace) & 15) return MAXK_ERR_ALIGN;
    if ((uintptr_t)g & 3) return MAXK_ERR_ALIGN;
    if ((k % 4 == 0) && ((uintptr_t)cbsr_sel & 3)) return MAXK_ERR_ALIGN;
    SchedWorkspace *ws = reinterpret_cast<SchedWorkspace *>(workspace);
    cudaError_t err;
    switch (k) {
        case 8: err = launch_bwd<8>(row_begin, row_end, indices, values, g, cbsr_sel, gs, n_rows, n_dst, n_edges, dim, (int)ld_g, k, row_div, ws, zero_fill, stream); break;
        case 16: err = launch_bwd<16>(row_begin, row_end, indices, values, g, cbsr_sel, gs, n_rows, n_dst, n_edges, dim, (int)ld_g, k, row_div, ws, zero_fill, stream); break;
        case 32: err = launch_bwd<32>(row_begin, row_end, indices, values, g, cbsr_sel, gs, n_rows, n_dst, n_edges, dim, (int)ld_g, k, row_div, ws, zero_fill, stream); break;
        case 64: err = launch_bwd<64>(row_begin, row_end, indices, values, g, cbsr_sel, gs, n_rows, n_dst, n_edges, dim, (int)ld_g, k, row_div, ws, zero_fill, stream); break;
        case 96: err = launch_bwd<96>(row_begin, row_end, indices, values, g, cbsr_sel, gs, n_rows, n_dst, n_edges, dim, (int)ld_g, k, row_div, ws, zero_fill, stream); break;
        case 128: err = launch_bwd<128>(row_begin, row_end, indices, values, g, cbsr_sel, gs, n_rows, n_dst, n_edges, dim, (int)ld_g, k, row_div, ws, zero_fill, stream); break;
        default: err = launch_bwd<0>(row_begin, row_end, indices, values, g, cbsr_sel, gs, n_rows, n_dst, n_edges, dim, (int)ld_g, k, row_div, ws, zero_fill, stream); break;
    }
    return status_from_cuda(err);
}

extern "C" int maxk_sspmm_backward(const int32_t *row_begin, const int32_t *row_end, const int32_t *indices,
                                   const float *values, const float *g, const uint8_t *cbsr_sel, float *gs,
                                   int64_t n_rows, int64_t n_dst, int64_t n_edges, int dim, int k,
                                   const float *row_div, void *workspace, size_t workspace_bytes,
                                   maxk_stream_t stream)
{
    return maxk_backward_strided(true, row_begin, row_end, indices, values, g, dim, cbsr_sel, gs, n_rows, n_dst, n_edges, dim,
                                 k, row_div, workspace, workspace_bytes, stream);
}

extern "C" int maxk_sspmm_backward_accumulate(const int32_t *row_begin, const int32_t *row_end,
                                              const int32_t *indices, const float *values, const float *g,
                                              const uint8_t *cbsr_sel, float *gs, int64_t n_rows, int64_t n_dst,
                                              int64_t n_edges, int dim, int k, const float *row_div, void *workspace,
                                              size_t workspace_bytes, maxk_stream_t stream)
{
    return maxk_backward_strided(false, row_begin, row_end, indices, values, g, dim, cbsr_sel, gs, n_rows, n_dst, n_edges, dim,
                                 k, row_div, workspace, workspace_bytes, stream);
}
