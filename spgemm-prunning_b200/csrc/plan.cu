// plan.cu -- the row plan of the slot-parallel forward SpGEMM, built on the GPU (sm_100a).
//
// The reference partitions rows into <= 64-edge warp segments on the host (kernels/generate_meta.py:30-48)
// and flushes every segment with 256 global atomics (kernels/spmm_maxk.cu:101-105).  Our forward owns whole
// rows instead and processes 8 of them in lockstep per warp (slots.cuh), so the partitioning metadata it
// needs is an ORDER: rows sorted by degree bucket, longest first (stable inside a bucket, so that
// neighbouring rows of a regular graph stay neighbours and share CSR sectors), cut into work items:
//   * rows with >= 4096 edges: one row per warp, edges dealt to the 8 slots ("shared" items);
//   * then groups of 8 rows of (nearly) equal degree, one per slot ("separate" items);
//   * when the graph has no mass of light rows to fill the end of the kernel with (a regular high-degree graph),
//     every warp gets the same whole number of groups and the rest of the rows is again run one row per warp, so
//     that the grid drains evenly whatever the problem size.
// Three small kernels (per-tile histogram, one-block scan, stable scatter), no host read-back: the plan can
// be built inside a CUDA graph and is a pure function of (row_begin, row_end, number of SMs).
#include "slots.cuh"

namespace maxk {

constexpr int kPlanBuckets = 256;
constexpr int kPlanTileRows = 1024;          // rows per warp tile
constexpr int kPlanThreads = 256;
constexpr int kPlanWarps = kPlanThreads / 32;

// degree -> bucket key, ascending key == descending degree.  8 sub-buckets per octave (rows of one
// bucket differ by < 12.5 % in length: the lockstep of a group wastes ~4 % on a log-normal degree
// distribution), exact for degrees < 16.  Boundaries at every power of two, so kPlanLongDeg and
// kPlanTailDeg are bucket boundaries.
__host__ __device__ inline int plan_key(int deg)
{
    if (deg <= 0) return kPlanBuckets - 1;
    int msb = 0;
    for (int d = deg; d > 1; d >>= 1) ++msb;
    const int fb = msb < 3 ? deg : 8 * (msb - 2) + ((deg >> (msb - 3)) & 7);
    return kPlanBuckets - 1 - fb;
}

__global__ void __launch_bounds__(kPlanThreads)
plan_hist_kernel(const int *__restrict__ row_begin, const int *__restrict__ row_end, int n_rows, int n_tiles,
                 int *__restrict__ hist)
{
    __shared__ int h[kPlanWarps][kPlanBuckets];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int tile = blockIdx.x * kPlanWarps + warp;
    for (int i = lane; i < kPlanBuckets; i += 32) h[warp][i] = 0;
    __syncwarp();
    if (tile < n_tiles) {
        const int r0 = tile * kPlanTileRows;
        for (int i = 0; i < kPlanTileRows / 32; ++i) {
            const int r = r0 + 32 * i + lane;
            if (r < n_rows) atomicAdd(&h[warp][plan_key(row_end[r] - row_begin[r])], 1);
        }
        __syncwarp();
        for (int i = lane; i < kPlanBuckets; i += 32) hist[(size_t)i * n_tiles + tile] = h[warp][i];
    }
}

// exclusive scan of hist (bucket-major) in place + plan header.  One block.
__global__ void __launch_bounds__(1024)
plan_scan_kernel(int *__restrict__ hist, int n_tiles, int n_rows, int total_warps, int *__restrict__ header)
{
    __shared__ int part[1024];
    const int total = kPlanBuckets * n_tiles;
    const int chunk = (total + 1023) / 1024;
    const int lo = min(total, (int)threadIdx.x * chunk), hi = min(total, lo + chunk);
    int s = 0;
    for (int i = lo; i < hi; ++i) s += hist[i];
    part[threadIdx.x] = s;
    __syncthreads();
    for (int off = 1; off < 1024; off <<= 1) {           // Hillis-Steele inclusive scan
        const int v = threadIdx.x >= off ? part[threadIdx.x - off] : 0;
        __syncthreads();
        part[threadIdx.x] += v;
        __syncthreads();
    }
    int run = part[threadIdx.x] - s;
    for (int i = lo; i < hi; ++i) {
        const int c = hist[i];
        hist[i] = run;
        run += c;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        // rows with deg >= D are the plan positions below the first bucket whose degrees are < D
        const int key_long = plan_key(kPlanLongDeg) + 1, key_tail = plan_key(kPlanTailDeg) + 1;
        const int n_long = hist[(size_t)key_long * n_tiles];
        const int p_tail = hist[(size_t)key_tail * n_tiles];       // rows with deg >= kPlanTailDeg
        // A group of 8 rows is indivisible and costs a warp 8 row-times, a single shared row 1.  When the graph
        // has no mass of light rows to even out the end of the kernel, every warp gets the same whole number of
        // groups and the remaining rows (>= 3 per warp) are dealt as single rows by the dynamic scheduler -- on a
        // small problem (one rank of an 8-GPU run: 12 rows per warp) a second group on a few warps would
        // otherwise set the kernel time (measured 0.43 ms instead of 0.31 ms for 1/8 of the Reddit shape).
        int posC = n_rows, posD = n_rows;
        if (n_rows - p_tail < 4 * total_warps && p_tail > n_long) {
            const int heavy = p_tail - n_long;                         // rows of 128 .. 4095 edges
            const int per_warp = heavy / total_warps;
            const int groups_per_warp = per_warp >= 3 ? (per_warp - 3) / kSS : 0;
            posC = n_long + min(groups_per_warp * total_warps, heavy / kSS) * kSS;
            posD = p_tail;
        }
        const int nA = n_long, nB = (posC - n_long + kSS - 1) / kSS, nC = posD - posC,
                  nD = (n_rows - posD + kSS - 1) / kSS;
        header[0] = kPlanMagic;
        header[1] = n_rows;
        header[2] = nA;
        header[3] = nB;
        header[4] = nC;
        header[5] = nD;
        header[6] = posC;
        header[7] = posD;
        header[8] = nA + nB + nC + nD;
        for (int i = 9; i < kPlanHeaderInts; ++i) header[i] = 0;
    }
}

__global__ void __launch_bounds__(kPlanThreads)
plan_scatter_kernel(const int *__restrict__ row_begin, const int *__restrict__ row_end, int n_rows, int n_tiles,
                    const int *__restrict__ scanned, int *__restrict__ p_row, int *__restrict__ p_beg,
                    int *__restrict__ p_end)
{
    __shared__ int base[kPlanWarps][kPlanBuckets];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int tile = blockIdx.x * kPlanWarps + warp;
    if (tile >= n_tiles) return;
    for (int i = lane; i < kPlanBuckets; i += 32) base[warp][i] = scanned[(size_t)i * n_tiles + tile];
    __syncwarp();
    const unsigned lt = (1u << lane) - 1u;
    const int r0 = tile * kPlanTileRows;
    for (int i = 0; i < kPlanTileRows / 32; ++i) {
        const int r = r0 + 32 * i + lane;
        int b = 0, e = 0, key = -1;
        if (r < n_rows) {
            b = row_begin[r];
            e = row_end[r];
            key = plan_key(e - b);
        }
        const unsigned peers = __match_any_sync(kFullMask, key);
        int pos = 0;
        if (key >= 0) pos = base[warp][key] + __popc(peers & lt);     // stable: lower lane == lower row first
        __syncwarp();
        if (key >= 0 && (peers & lt) == 0) base[warp][key] += __popc(peers);
        __syncwarp();
        if (key >= 0) {
            p_row[pos] = r;
            p_beg[pos] = b;
            p_end[pos] = e;
        }
    }
}

}  // namespace maxk

using namespace maxk;

extern "C" size_t maxk_plan_bytes(int64_t n_rows)
{
    if (n_rows < 0) n_rows = 0;
    return sizeof(int) * ((size_t)kPlanHeaderInts + 3 * (size_t)plan_pad_rows(n_rows) + 2 * kPlanTicketSlots);
}

extern "C" size_t maxk_plan_workspace_bytes(int64_t n_rows)
{
    if (n_rows < 0) n_rows = 0;
    const size_t n_tiles = ((size_t)n_rows + kPlanTileRows - 1) / kPlanTileRows;
    return sizeof(int) * kPlanBuckets * (n_tiles > 0 ? n_tiles : 1) + 16;
}

// warps of the forward's persistent grid: 2 CTAs x 8 warps per SM (spgemm_fwd.cu)
int plan_total_warps()
{
    return device_sm_count() * 16;
}

extern "C" int maxk_plan_build(const int32_t *row_begin, const int32_t *row_end, int64_t n_rows, void *plan,
                               size_t plan_bytes, void *workspace, size_t workspace_bytes, maxk_stream_t stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    if (n_rows < 0 || n_rows > INT32_MAX - 64) return MAXK_ERR_SIZE;
    if (!plan || !workspace) return MAXK_ERR_NULL;
    if (n_rows > 0 && (!row_begin || !row_end)) return MAXK_ERR_NULL;
    if (plan_bytes < maxk_plan_bytes(n_rows) || workspace_bytes < maxk_plan_workspace_bytes(n_rows)) return MAXK_ERR_WORKSPACE;
    if (((uintptr_t)plan | (uintptr_t)workspace) & 15) return MAXK_ERR_ALIGN;
    int *header = reinterpret_cast<int *>(plan);
    int *hist = reinterpret_cast<int *>(workspace);
    const int n_tiles = (int)((n_rows + kPlanTileRows - 1) / kPlanTileRows);
    const int64_t n_pad = plan_pad_rows(n_rows);
    int *p_row = header + kPlanHeaderInts, *p_beg = p_row + n_pad, *p_end = p_beg + n_pad;
    const int tiles = n_tiles > 0 ? n_tiles : 1;
    {   // scheduler tickets of the launches that will use this plan
        cudaError_t err = cudaMemsetAsync(p_end + n_pad, 0, sizeof(int) * 2 * kPlanTicketSlots, stream);
        if (err != cudaSuccess) return status_from_cuda(err);
    }
    if (n_tiles == 0) {
        cudaError_t err = cudaMemsetAsync(hist, 0, sizeof(int) * kPlanBuckets, stream);
        if (err != cudaSuccess) return status_from_cuda(err);
    } else {
        plan_hist_kernel<<<(n_tiles + kPlanWarps - 1) / kPlanWarps, kPlanThreads, 0, stream>>>(row_begin, row_end, (int)n_rows,
                                                                                               n_tiles, hist);
    }
    plan_scan_kernel<<<1, 1024, 0, stream>>>(hist, tiles, (int)n_rows, plan_total_warps(), header);
    if (n_tiles > 0)
        plan_scatter_kernel<<<(n_tiles + kPlanWarps - 1) / kPlanWarps, kPlanThreads, 0, stream>>>(
            row_begin, row_end, (int)n_rows, n_tiles, hist, p_row, p_beg, p_end);
    return status_from_cuda(cudaGetLastError());
}
