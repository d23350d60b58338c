// meta.cu -- warp/row-block partition metadata (warp4) built on the GPU, and C-ABI housekeeping.
//
// Replaces the per-row CPython loop + file round trip of the reference
// (kernels/generate_meta.py:30-48 -> w12_nz64_warp_4/<graph>.warp4 -> load_warp4_metadata,
// cuda_kernel_bindings.cpp:287-317).  Same output contract: int32 quads
// (row, loc, len <= max_nz, 0) for every non-empty CSR row, in row order.
//   scan : segs[r] = ceil(deg[r]/max_nz), three-phase exclusive scan -> seg_offsets[n_rows+1]
//   fill : one thread per QUAD (binary search of its row in seg_offsets): balanced for any
//          degree skew, 16-byte coalesced stores.
#include "maxk_common.cuh"

namespace maxk {

constexpr int kScanThreads = 256;
constexpr int kScanItems = 4;                           // rows per thread
constexpr int kScanTile = kScanThreads * kScanItems;    // rows per block

__device__ __forceinline__ int block_excl_scan(int v, int &total)
{
    __shared__ int s_warp[kScanThreads / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int n = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += n;
    }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        int w = lane < kScanThreads / 32 ? s_warp[lane] : 0;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int n = __shfl_up_sync(0xffffffffu, w, d);
            if (lane >= d) w += n;
        }
        if (lane < kScanThreads / 32) s_warp[lane] = w;  // inclusive
    }
    __syncthreads();
    const int warp_off = warp > 0 ? s_warp[warp - 1] : 0;
    total = s_warp[kScanThreads / 32 - 1];
    __syncthreads();
    return warp_off + inc - v;
}

__device__ __forceinline__ int segs_of(const int *indptr, int64_t r, int max_nz)
{
    const int deg = indptr[r + 1] - indptr[r];
    return deg > 0 ? (deg + max_nz - 1) / max_nz : 0;
}

// phase 1: per-tile local exclusive scan + tile totals
__global__ void __launch_bounds__(kScanThreads)
warp4_scan_tiles(const int *__restrict__ indptr, int64_t n_rows, int max_nz, int *__restrict__ seg_offsets,
                 int *__restrict__ tile_totals)
{
    const int64_t base = (int64_t)blockIdx.x * kScanTile + (int64_t)threadIdx.x * kScanItems;
    int c[kScanItems];
    int sum = 0;
#pragma unroll
    for (int i = 0; i < kScanItems; ++i) {
        c[i] = (base + i < n_rows) ? segs_of(indptr, base + i, max_nz) : 0;
        sum += c[i];
    }
    int total;
    int off = block_excl_scan(sum, total);
#pragma unroll
    for (int i = 0; i < kScanItems; ++i) {
        if (base + i < n_rows) seg_offsets[base + i] = off;
        off += c[i];
    }
    if (threadIdx.x == 0) tile_totals[blockIdx.x] = total;
}

// phase 2: one block scans the tile totals in place (exclusive); writes the grand total.
__global__ void __launch_bounds__(kScanThreads)
warp4_scan_totals(int *__restrict__ tile_totals, int n_tiles, int *__restrict__ grand_total)
{
    int carry = 0;
    for (int base = 0; base < n_tiles; base += kScanThreads) {
        const int i = base + threadIdx.x;
        const int v = i < n_tiles ? tile_totals[i] : 0;
        int total;
        const int off = block_excl_scan(v, total);
        if (i < n_tiles) tile_totals[i] = carry + off;
        carry += total;
    }
    if (threadIdx.x == 0) *grand_total = carry;
}

// phase 3: add tile offsets
__global__ void __launch_bounds__(kScanThreads)
warp4_scan_apply(int64_t n_rows, int *__restrict__ seg_offsets, const int *__restrict__ tile_totals)
{
    const int64_t base = (int64_t)blockIdx.x * kScanTile + (int64_t)threadIdx.x * kScanItems;
    const int add = tile_totals[blockIdx.x];
#pragma unroll
    for (int i = 0; i < kScanItems; ++i)
        if (base + i < n_rows) seg_offsets[base + i] += add;
}

__global__ void __launch_bounds__(256)
warp4_fill_kernel(const int *__restrict__ indptr, const int *__restrict__ seg_offsets, int64_t n_rows, int max_nz,
                  int4 *__restrict__ warp4)
{
    const int n_quads = seg_offsets[n_rows];
    for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < n_quads; q += gridDim.x * blockDim.x) {
        // last row r with seg_offsets[r] <= q (rows with zero segments share an offset with
        // their successor, so take the upper bound and step back)
        int64_t lo = 0, hi = n_rows;  // invariant: seg_offsets[lo] <= q < seg_offsets[hi]
        while (hi - lo > 1) {
            const int64_t mid = (lo + hi) >> 1;
            if (seg_offsets[mid] <= q) lo = mid; else hi = mid;
        }
        const int r = (int)lo;
        const int j = q - seg_offsets[r];
        const int beg = indptr[r], end = indptr[r + 1];
        const int loc = beg + j * max_nz;
        warp4[q] = make_int4(r, loc, min(max_nz, end - loc), 0);
    }
}

__global__ void __launch_bounds__(256)
warp4_to_rows_kernel(const int4 *__restrict__ warp4, int64_t n_quads, int n_rows, int *__restrict__ row_begin,
                     int *__restrict__ row_end)
{
    for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < n_quads; q += (int64_t)gridDim.x * blockDim.x) {
        const int4 w = warp4[q];
        if (w.x < 0 || w.x >= n_rows) continue;  // malformed quad: ignore rather than write out of bounds
        const bool first = (q == 0) || (warp4[q - 1].x != w.x);
        const bool last = (q == n_quads - 1) || (warp4[q + 1].x != w.x);
        if (first) row_begin[w.x] = w.y;
        if (last) row_end[w.x] = w.y + w.z;
    }
}

}  // namespace maxk

using namespace maxk;

extern "C" int maxk_abi_version(void) { return 1; }

extern "C" const char *maxk_status_string(int status)
{
    switch (status) {
        case MAXK_OK: return "ok";
        case MAXK_ERR_BAD_K: return "k must satisfy 1 <= k <= dim";
        case MAXK_ERR_BAD_DIM: return "dim must satisfy 1 <= dim <= 256 (uint8 column selectors)";
        case MAXK_ERR_NULL: return "a required pointer is NULL";
        case MAXK_ERR_WORKSPACE: return "workspace too small";
        case MAXK_ERR_ALIGN: return "pointer not sufficiently aligned (16 B for fp32 matrices, 4 B for selectors)";
        case MAXK_ERR_SIZE: return "negative or too large size";
        default: return status > 0 ? cudaGetErrorString((cudaError_t)status) : "unknown maxk status";
    }
}

extern "C" size_t maxk_warp4_workspace_bytes(int64_t n_rows)
{
    if (n_rows < 0) n_rows = 0;
    const int64_t tiles = (n_rows + kScanTile - 1) / kScanTile;
    return sizeof(int) * (size_t)(tiles + 1);
}

extern "C" int maxk_warp4_scan(const int32_t *indptr, int64_t n_rows, int max_nz, int32_t *seg_offsets,
                               void *workspace, size_t workspace_bytes, maxk_stream_t stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    if (n_rows < 0 || n_rows > INT32_MAX || max_nz < 1) return MAXK_ERR_SIZE;
    if (!seg_offsets) return MAXK_ERR_NULL;
    if (n_rows == 0) return status_from_cuda(cudaMemsetAsync(seg_offsets, 0, sizeof(int), stream));
    if (!indptr || !workspace) return MAXK_ERR_NULL;
    if (workspace_bytes < maxk_warp4_workspace_bytes(n_rows)) return MAXK_ERR_WORKSPACE;
    int *tile_totals = reinterpret_cast<int *>(workspace);
    const int tiles = (int)((n_rows + kScanTile - 1) / kScanTile);
    warp4_scan_tiles<<<tiles, kScanThreads, 0, stream>>>(indptr, n_rows, max_nz, seg_offsets, tile_totals);
    warp4_scan_totals<<<1, kScanThreads, 0, stream>>>(tile_totals, tiles, seg_offsets + n_rows);
    warp4_scan_apply<<<tiles, kScanThreads, 0, stream>>>(n_rows, seg_offsets, tile_totals);
    return status_from_cuda(cudaGetLastError());
}

extern "C" int maxk_warp4_fill(const int32_t *indptr, const int32_t *seg_offsets, int64_t n_rows, int max_nz,
                               int32_t *warp4, maxk_stream_t stream)
{
    if (n_rows < 0 || n_rows > INT32_MAX || max_nz < 1) return MAXK_ERR_SIZE;
    if (n_rows == 0) return MAXK_OK;
    if (!indptr || !seg_offsets || !warp4) return MAXK_ERR_NULL;
    if ((uintptr_t)warp4 & 15) return MAXK_ERR_ALIGN;
    int dev = 0, sms = kNumSMsB200;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    warp4_fill_kernel<<<sms * 8, 256, 0, (cudaStream_t)stream>>>(indptr, seg_offsets, n_rows, max_nz,
                                                                reinterpret_cast<int4 *>(warp4));
    return status_from_cuda(cudaGetLastError());
}

extern "C" int maxk_warp4_to_rows(const int32_t *warp4, int64_t n_quads, int64_t n_rows, int32_t *row_begin,
                                  int32_t *row_end, maxk_stream_t stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    if (n_rows < 0 || n_rows > INT32_MAX || n_quads < 0) return MAXK_ERR_SIZE;
    if (n_rows == 0) return MAXK_OK;
    if (!row_begin || !row_end) return MAXK_ERR_NULL;
    cudaError_t err = cudaMemsetAsync(row_begin, 0, sizeof(int) * (size_t)n_rows, stream);
    if (err == cudaSuccess) err = cudaMemsetAsync(row_end, 0, sizeof(int) * (size_t)n_rows, stream);
    if (err != cudaSuccess) return status_from_cuda(err);
    if (n_quads == 0) return MAXK_OK;
    if (!warp4) return MAXK_ERR_NULL;
    if ((uintptr_t)warp4 & 15) return MAXK_ERR_ALIGN;
    int dev = 0, sms = kNumSMsB200;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    warp4_to_rows_kernel<<<sms * 8, 256, 0, stream>>>(reinterpret_cast<const int4 *>(warp4), n_quads, (int)n_rows,
                                                     row_begin, row_end);
    return status_from_cuda(cudaGetLastError());
}
