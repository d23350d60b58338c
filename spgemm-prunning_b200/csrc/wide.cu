// wide.cu -- feature widths above 256: CBSR with uint16 column selectors (sm_100a).
//
// The reference addresses columns with uint8 (`topk_indices.to(torch.uint8)`, maxk_spgemm_function.py:57,
// cuda_kernel_bindings.cpp:70 hard-wires 256 columns), so its own Yelp script (scripts_train/yelp_maxk.sh:16,
// hidden 384) silently wraps column ids.  This file lifts the limit to dim <= 1024:
//   * maxk_topk_cbsr16: exact row-wise top-k of x[N, dim <= 1024] -> k fp32 values + k uint16 columns, column
//     ascending (ties: lowest column, NaN largest, -0 == +0, as the 256-column kernels);
//   * maxk_spgemm_forward16 / maxk_sspmm_backward16: the 256-column kernels run once per 256-column BLOCK of the
//     feature matrix, in place on the strided [N, dim] output / gradient.  For block p a small kernel rewrites the
//     CBSR rows into block-local uint8 form: an entry inside the block keeps its value and gets column c - 256 p;
//     an entry outside the block becomes a zero-valued entry on a column of the block that the row does not use
//     (the forward's accumulate needs the columns of a row to be distinct).  Entry positions are preserved, so the
//     backward's per-block sampled gradients are merged by picking, for every entry, the block it lives in.
// Cost: ceil(dim / 256) passes over the edges (parity first; a native wide accumulator is the follow-up).
#include "maxk_common.cuh"

int maxk_forward_strided(const void *plan, const int32_t *indices, const float *values, const float *cbsr_val,
                         const uint8_t *cbsr_sel, float *out, int64_t ld_out, int64_t n_rows, int64_t n_edges, int dim,
                         int k, const float *row_div, cudaStream_t stream);
int maxk_backward_strided(bool zero_fill, const int32_t *row_begin, const int32_t *row_end, const int32_t *indices,
                          const float *values, const float *g, int64_t ld_g, const uint8_t *cbsr_sel, float *gs,
                          int64_t n_rows, int64_t n_dst, int64_t n_edges, int dim, int k, const float *row_div,
                          void *workspace, size_t workspace_bytes, maxk_stream_t stream_);

namespace maxk {

constexpr int kWideMaxDim = 1024;
constexpr int kWideThreads = 256;
constexpr int kWideWarps = kWideThreads / 32;
constexpr unsigned kFullW = 0xffffffffu;

__device__ __forceinline__ uint32_t wide_key(float v)
{
    const uint32_t b = __float_as_uint(__fadd_rn(v, 0.0f));      // -0 -> +0, every NaN -> the canonical quiet NaN
    return b ^ ((uint32_t)((int32_t)b >> 31) | 0x80000000u);
}

// One warp per row; lane l owns columns l, l + 32, ... (NPL = ceil(dim / 32) <= 32 keys per lane, coalesced loads).
// Threshold = the k-th largest key, built bit by bit (32 warp-wide counts).
template <int NPL>
__global__ void __launch_bounds__(kWideThreads)
topk16_kernel(const float *__restrict__ x, int64_t n_rows, int dim, int k, float *__restrict__ out_val,
              uint16_t *__restrict__ out_sel, float *__restrict__ masked)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned lt = (1u << lane) - 1u;
    const int64_t warps_total = (int64_t)gridDim.x * kWideWarps;
    for (int64_t r = (int64_t)blockIdx.x * kWideWarps + warp; r < n_rows; r += warps_total) {
        const float *row = x + r * dim;
        float v[NPL];
        uint32_t key[NPL];
#pragma unroll
        for (int j = 0; j < NPL; ++j) {
            const int c = lane + 32 * j;
            v[j] = c < dim ? row[c] : 0.f;
            key[j] = c < dim ? wide_key(v[j]) : 0u;              // pad: below every real key (those are >= 0x007fffff)
        }
        uint32_t T = 0u;
        for (int bit = 31; bit >= 0; --bit) {
            const uint32_t cand = T | (1u << bit);
            int c = 0;
#pragma unroll
            for (int j = 0; j < NPL; ++j) c += key[j] >= cand ? 1 : 0;
            if (__reduce_add_sync(kFullW, c) >= k) T = cand;
        }
        // key > T always; key == T in column order (j major, lane minor) until k entries are taken
        int gt_total = 0;
#pragma unroll
        for (int j = 0; j < NPL; ++j) gt_total += __popc(__ballot_sync(kFullW, key[j] > T));
        const int need_eq = k - gt_total;
        int eq_seen = 0, pos_base = 0;
        float *mrow = masked ? masked + r * dim : nullptr;
#pragma unroll
        for (int j = 0; j < NPL; ++j) {
            const bool is_eq = key[j] == T;
            const unsigned b_eq = __ballot_sync(kFullW, is_eq);
            const bool take = key[j] > T || (is_eq && eq_seen + __popc(b_eq & lt) < need_eq);
            eq_seen += __popc(b_eq);
            const unsigned b_sel = __ballot_sync(kFullW, take);
            const int c = lane + 32 * j;
            if (take) {
                const int64_t o = r * k + pos_base + __popc(b_sel & lt);
                out_val[o] = v[j];
                out_sel[o] = (uint16_t)c;
            }
            if (mrow && c < dim) mrow[c] = take ? v[j] : 0.f;
            pos_base += __popc(b_sel);
        }
    }
}

// Block-local uint8 CBSR of column block p (see the file header).  One warp per row.
__global__ void __launch_bounds__(kWideThreads)
cbsr16_split_kernel(const float *__restrict__ vals, const uint16_t *__restrict__ sel16, int64_t n_rows, int k, int block,
                    float *__restrict__ out_val, uint8_t *__restrict__ out_sel)
{
    __shared__ uint32_t s_used[kWideWarps][8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t warps_total = (int64_t)gridDim.x * kWideWarps;
    for (int64_t r = (int64_t)blockIdx.x * kWideWarps + warp; r < n_rows; r += warps_total) {
        if (lane < 8) s_used[warp][lane] = 0u;
        __syncwarp();
        for (int l = lane; l < k; l += 32) {
            const int c = sel16[r * k + l];
            if ((c >> 8) == block) atomicOr(&s_used[warp][(c & 255) >> 5], 1u << (c & 31));
        }
        __syncwarp();
        int outside_before = 0;                                 // entries outside the block in positions < l
        for (int l0 = 0; l0 < k; l0 += 32) {
            const int l = l0 + lane;
            const int c = l < k ? sel16[r * k + l] : (block << 8);
            const bool outside = l < k && (c >> 8) != block;
            const unsigned b_out = __ballot_sync(kFullW, outside);
            if (l < k) {
                float v = vals ? vals[r * k + l] : 0.f;
                int col = c & 255;
                if (outside) {                                  // the j-th unused column of the block
                    int j = outside_before + __popc(b_out & ((1u << lane) - 1u));
                    col = 0;
                    for (int w = 0; w < 8; ++w) {
                        const uint32_t free_bits = ~s_used[warp][w];
                        const int cnt = __popc(free_bits);
                        if (j < cnt) {
                            uint32_t m = free_bits;
                            for (int i = 0; i < j; ++i) m &= m - 1;      // drop the j lowest set bits
                            col = 32 * w + (__ffs(m) - 1);
                            break;
                        }
                        j -= cnt;
                    }
                    v = 0.f;
                }
                if (out_val) out_val[r * k + l] = v;
                out_sel[r * k + l] = (uint8_t)col;
            }
            outside_before += __popc(b_out);
        }
        __syncwarp();
    }
}

// gs[r, l] = parts[block of entry l][r, l]
__global__ void __launch_bounds__(kWideThreads)
cbsr16_merge_kernel(const float *__restrict__ parts, const uint16_t *__restrict__ sel16, int64_t n_entries,
                    float *__restrict__ gs)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_entries; i += stride)
        gs[i] = parts[(int64_t)(sel16[i] >> 8) * n_entries + i];
}

static int wide_grid(int64_t work_items, int per_block)
{
    const int64_t need = (work_items + per_block - 1) / per_block;
    const int64_t cap = (int64_t)device_sm_count() * 8;
    return (int)(need < 1 ? 1 : (need < cap ? need : cap));
}

static size_t align256(size_t v) { return (v + 255) / 256 * 256; }

}  // namespace maxk

using namespace maxk;

extern "C" size_t maxk_plan_bytes(int64_t n_rows);
extern "C" size_t maxk_spgemm_workspace_bytes(int64_t n_rows);

extern "C" int maxk_topk_cbsr16(const float *x, int64_t n_rows, int dim, int k, float *cbsr_val, uint16_t *cbsr_sel,
                                float *masked, maxk_stream_t stream)
{
    if (dim < 1 || dim > kWideMaxDim) return MAXK_ERR_BAD_DIM;
    if (k < 1 || k > dim || k > kAccDim) return MAXK_ERR_BAD_K;
    if (n_rows < 0) return MAXK_ERR_SIZE;
    if (n_rows == 0) return MAXK_OK;
    if (!x || !cbsr_val || !cbsr_sel) return MAXK_ERR_NULL;
    cudaStream_t st = (cudaStream_t)stream;
    const int grid = wide_grid(n_rows, kWideWarps);
    const int npl = (dim + 31) / 32;
    if (npl <= 8) topk16_kernel<8><<<grid, kWideThreads, 0, st>>>(x, n_rows, dim, k, cbsr_val, cbsr_sel, masked);
    else if (npl <= 16) topk16_kernel<16><<<grid, kWideThreads, 0, st>>>(x, n_rows, dim, k, cbsr_val, cbsr_sel, masked);
    else if (npl <= 24) topk16_kernel<24><<<grid, kWideThreads, 0, st>>>(x, n_rows, dim, k, cbsr_val, cbsr_sel, masked);
    else topk16_kernel<32><<<grid, kWideThreads, 0, st>>>(x, n_rows, dim, k, cbsr_val, cbsr_sel, masked);
    return status_from_cuda(cudaGetLastError());
}

/* scratch of the two wide operators: one block-local CBSR (values + uint8 selectors), for the backward one sampled
 * gradient per block, plus the scratch of the 256-column kernels */
extern "C" size_t maxk_wide_workspace_bytes(int64_t n_rows, int64_t n_src, int dim, int k)
{
    if (n_rows < 0) n_rows = 0;
    if (n_src < 0) n_src = 0;
    const size_t blocks = (size_t)((dim > 0 ? dim : 1) + kAccDim - 1) / kAccDim;
    const size_t entries = (size_t)n_src * (size_t)(k > 0 ? k : 1);
    return align256(entries * 4) + align256(entries) + align256(blocks * entries * 4) + align256(maxk_spgemm_workspace_bytes(n_rows));
}

extern "C" int maxk_spgemm_forward16(const void *plan, const int32_t *indices, const float *values, const float *cbsr_val,
                                     const uint16_t *cbsr_sel, float *out, int64_t n_rows, int64_t n_src, int64_t n_edges,
                                     int dim, int k, const float *row_div, void *workspace, size_t workspace_bytes,
                                     maxk_stream_t stream)
{
    if (dim < 1 || dim > kWideMaxDim) return MAXK_ERR_BAD_DIM;
    if (k < 1 || k > kAccDim) return MAXK_ERR_BAD_K;
    if (n_rows < 0 || n_src < 0 || n_edges < 0) return MAXK_ERR_SIZE;
    if (n_rows == 0) return MAXK_OK;
    if (!plan || !out || !workspace) return MAXK_ERR_NULL;
    if (n_src > 0 && (!cbsr_val || !cbsr_sel)) return MAXK_ERR_NULL;
    if (workspace_bytes < maxk_wide_workspace_bytes(n_rows, n_src, dim, k)) return MAXK_ERR_WORKSPACE;
    if ((uintptr_t)workspace & 255) return MAXK_ERR_ALIGN;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t entries = (size_t)n_src * k;
    unsigned char *base = reinterpret_cast<unsigned char *>(workspace);
    float *bval = reinterpret_cast<float *>(base);
    uint8_t *bsel = base + align256(entries * 4);
    const int blocks = (dim + kAccDim - 1) / kAccDim;
    for (int p = 0; p < blocks; ++p) {
        if (n_src > 0) {
            cbsr16_split_kernel<<<wide_grid(n_src, kWideWarps), kWideThreads, 0, st>>>(cbsr_val, cbsr_sel, n_src, k, p, bval, bsel);
            const cudaError_t e = cudaGetLastError();
            if (e != cudaSuccess) return status_from_cuda(e);
        }
        const int width = dim - kAccDim * p < kAccDim ? dim - kAccDim * p : kAccDim;
        const int st_code = maxk_forward_strided(plan, indices, values, bval, bsel, out + (size_t)kAccDim * p, dim, n_rows,
                                                 n_edges, width, k, row_div, st);
        if (st_code != MAXK_OK) return st_code;
    }
    return MAXK_OK;
}

extern "C" int maxk_sspmm_backward16(const int32_t *row_begin, const int32_t *row_end, const int32_t *indices,
                                     const float *values, const float *g, const uint16_t *cbsr_sel, float *gs,
                                     int64_t n_rows, int64_t n_dst, int64_t n_edges, int dim, int k, const float *row_div,
                                     void *workspace, size_t workspace_bytes, maxk_stream_t stream)
{
    if (dim < 1 || dim > kWideMaxDim) return MAXK_ERR_BAD_DIM;
    if (k < 1 || k > kAccDim) return MAXK_ERR_BAD_K;
    if (n_rows < 0 || n_dst < 0 || n_edges < 0) return MAXK_ERR_SIZE;
    if (n_dst == 0) return MAXK_OK;
    if (!gs || !cbsr_sel || !workspace) return MAXK_ERR_NULL;
    if (workspace_bytes < maxk_wide_workspace_bytes(n_rows, n_dst, dim, k)) return MAXK_ERR_WORKSPACE;
    if ((uintptr_t)workspace & 255) return MAXK_ERR_ALIGN;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t entries = (size_t)n_dst * k;
    const int blocks = (dim + kAccDim - 1) / kAccDim;
    unsigned char *base = reinterpret_cast<unsigned char *>(workspace);
    uint8_t *bsel = base + align256(entries * 4);                                    // the backward needs no values
    float *parts = reinterpret_cast<float *>(base + align256(entries * 4) + align256(entries));
    void *inner = base + align256(entries * 4) + align256(entries) + align256((size_t)blocks * entries * 4);
    const size_t inner_bytes = maxk_spgemm_workspace_bytes(n_rows);
    for (int p = 0; p < blocks; ++p) {
        cbsr16_split_kernel<<<wide_grid(n_dst, kWideWarps), kWideThreads, 0, st>>>(nullptr, cbsr_sel, n_dst, k, p, nullptr, bsel);
        const cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return status_from_cuda(e);
        const int width = dim - kAccDim * p < kAccDim ? dim - kAccDim * p : kAccDim;
        const int st_code = maxk_backward_strided(true, row_begin, row_end, indices, values, g + (size_t)kAccDim * p, dim, bsel,
                                                  parts + (size_t)p * entries, n_rows, n_dst, n_edges, width, k, row_div,
                                                  inner, inner_bytes, stream);
        if (st_code != MAXK_OK) return st_code;
    }
    cbsr16_merge_kernel<<<wide_grid((int64_t)entries, kWideThreads), kWideThreads, 0, st>>>(parts, cbsr_sel, (int64_t)entries, gs);
    return status_from_cuda(cudaGetLastError());
}
