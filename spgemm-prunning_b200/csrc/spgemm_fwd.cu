// spgemm_fwd.cu -- forward row-wise-product SpGEMM, CSR adjacency x CBSR features (sm_100a),
// slot-parallel (slots.cuh) over the row plan (plan.cu).
//
// Replaces spmm_kernel_opt2_sparse_v3 (reference kernels/spmm_maxk.cu:17-106):
//   out[r, sel[c,l]] += val[e] * data[c,l]     for every edge e = (r, c), l < k
//   * rows are OWNED: a row is reduced by one slot (or, for long rows, by the 8 slots of one warp) into
//     shared-memory accumulators and written exactly once.  No global atomics, no output memset, a fixed
//     reduction order (the reference flushes 256 global atomics per <= 64-edge segment,
//     spmm_maxk.cu:101-105, into a torch::zeros output, cuda_kernel_bindings.cpp:71);
//   * 8 edges per warp instruction, each gathered as one 32-byte value load + one 8-byte selector load per
//     lane (k = 32): every 128-byte line of a neighbour's CBSR row is touched by exactly one instruction;
//   * the degree normalisation the reference does in a separate PyTorch pass (maxk_spgemm_function.py:86)
//     is fused into the epilogue;
//   * work items are handed to a persistent grid in plan order (longest first) by one atomic ticket counter that
//     the last warp resets: no memset, no workspace, no second "long row" kernel.
#include "slots.cuh"
#include <atomic>

namespace maxk {

constexpr int kFsThreads = 256;
constexpr int kFsWarps = kFsThreads / 32;

// steps (x 8 edges) gathered before the first accumulate; MINB = CTAs per SM the kernel is compiled for
template <int K, int MINB> struct FwdTune {
    static constexpr int UNROLL = MINB >= 3 ? (K <= 32 ? 2 : 1) : (K <= 32 ? 4 : K <= 64 ? 2 : 1);
};

// one step of one slot: acc[col] += w * v for the lane's entries
template <int K>
__device__ __forceinline__ void accumulate_entries(const SlotEntries<K> &en, float w, float *acc_q)
{
    constexpr int EPL = SlotEntries<K>::EPL;
    constexpr int B = EPL < 8 ? EPL : 8;      // read B, then write B: the columns of one CBSR row are distinct
#pragma unroll
    for (int i0 = 0; i0 < EPL; i0 += B) {
        int o[B];
        float a[B];
#pragma unroll
        for (int i = 0; i < B; ++i) {
            o[i] = slot_word(en.col(i0 + i));
            a[i] = acc_q[o[i]];
        }
#pragma unroll
        for (int i = 0; i < B; ++i) acc_q[o[i]] = fmaf(w, en.v[i0 + i], a[i]);
    }
}

// any k in [1, 256] (K == 0): lane t of a slot walks entries t, t+4, ... with scalar loads
__device__ __forceinline__ void accumulate_any_k(const float *__restrict__ cval, const uint8_t *__restrict__ csel,
                                                 size_t row_off, int k, int t, float w, float *acc_q)
{
    for (int l = t; l < k; l += kSL) {
        const int o = slot_word(__ldg(csel + row_off + l));
        acc_q[o] = fmaf(w, __ldg(cval + row_off + l), acc_q[o]);
    }
}

template <int K, int MINB>
__global__ void __launch_bounds__(kFsThreads, MINB)
spgemm_fwd_slots_kernel(const int *__restrict__ plan, const int *__restrict__ idx, const float *__restrict__ val,
                        const float *__restrict__ cval, const uint8_t *__restrict__ csel, float *__restrict__ out,
                        int n_edges, int dim, int ld_out, int k, const float *__restrict__ row_div, int ticket_slot)
{
    const bool vec_out = dim == kAccDim && (ld_out & 7) == 0;     // 32-byte row stores (out is 32-byte aligned)
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = lane_id(), warp = threadIdx.x >> 5;
    float *acc = reinterpret_cast<float *>(smem_raw + (size_t)warp * kSlotWarpBytes);
    int2 *cw = reinterpret_cast<int2 *>(acc + kSlotCopyWords);
    float4 *acc4 = reinterpret_cast<float4 *>(acc);
    for (int i = lane; i < kSlotCopyWords / 4; i += 32) acc4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    __syncwarp();

    const PlanView pv = load_plan(plan);
    const int q = lane / kSL, t = lane % kSL;
    float *acc_q = acc + kSL * q;
    const uint64_t keep = policy_evict_last();        // CBSR rows are re-used ~degree times: keep them in L2
    const int total_warps = gridDim.x * kFsWarps;
    int *ticket = pv.tickets + 2 * ticket_slot;
    // The first item of a warp is its own index (the heaviest items, one each); later items come from the ticket
    // counter in plan order.  A warp holds ONE claimed item ahead while it works on a long item and two while it
    // works on short ones (their descriptors have to be in flight a few microseconds ahead): claiming two heavy
    // items ahead let the first warps hoard the groups of a small problem (1/8 of the Reddit shape: 0.43 ms
    // instead of 0.31 ms).
    const int item = blockIdx.x * kFsWarps + warp;
    if (item >= pv.n_items) {
        leave_scheduler(ticket, lane, total_warps);
        return;
    }
    Item it_cur = decode_item(pv, item);
    Desc d_cur = load_desc(pv, it_cur, lane);
    bool exhausted = false;
    bool has_nxt = false, has_nxt2 = false;
    Desc d_nxt = d_cur, d_nxt2 = d_cur;
    int sh_nxt = 0, sh_nxt2 = 0;
    {
        const int i = total_warps + grab_item(ticket, lane);
        has_nxt = i < pv.n_items;
        exhausted = !has_nxt;
        if (has_nxt) {
            const Item it = decode_item(pv, i);
            d_nxt = load_desc(pv, it, lane);
            sh_nxt = it.shared;
        }
    }
    int pc[kSS / 2];
    float pw[kSS / 2];
    WindowMap wm = make_window_map(d_cur, it_cur.shared, lane);
    fetch_window(wm, idx, val, 0, n_edges, pc, pw);

    for (;;) {
        const SlotView sv = make_slot_view(d_cur, it_cur.shared, lane);
        if (has_nxt && !has_nxt2 && !exhausted && sv.steps < kAlignedSteps) {     // short item: a second one ahead
            const int i = total_warps + grab_item(ticket, lane);
            has_nxt2 = i < pv.n_items;
            exhausted = !has_nxt2;
            if (has_nxt2) {
                const Item it = decode_item(pv, i);
                d_nxt2 = load_desc(pv, it, lane);
                sh_nxt2 = it.shared;
            }
        }
        // the registers pc/pw always hold the next window to park: window `parked` of this item, or, once all
        // of them are parked, window 0 of the next item
        const int n_win = wm.n_win, lead = wm.lead;
        int parked = 0;
        bool next_fetched = false;
        int j_blk = 0;
        for (int s0 = 0; s0 < sv.steps; s0 += kSW, ++j_blk) {
            while (parked <= j_blk + lead && parked < n_win) {
                __syncwarp();
                park_window(wm, cw, parked, pc, pw);
                ++parked;
                if (parked < n_win) {
                    fetch_window(wm, idx, val, parked, n_edges, pc, pw);
                } else if (has_nxt) {
                    const WindowMap wn = make_window_map(d_nxt, sh_nxt, lane);
                    fetch_window(wn, idx, val, 0, n_edges, pc, pw);
                    next_fetched = true;
                }
            }
            __syncwarp();
            const int nj = min(kSW, sv.steps - s0);
            if constexpr (K != 0) {
                constexpr int U = FwdTune<K, MINB>::UNROLL;
                for (int j = 0; j < nj; j += U) {
                    SlotEntries<K> en[U];
                    float w[U];
                    bool ok[U];
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        const int2 e = read_window(sv, cw, s0 + j + u);
                        w[u] = __int_as_float(e.y);
                        ok[u] = s0 + j + u < sv.cnt;
                        if (ok[u]) {
                            const size_t row_off = (size_t)e.x * K;
                            en[u].load_val(cval, row_off, t, keep);
                            en[u].load_sel(csel, row_off, t, keep);
                        }
                    }
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        if (ok[u]) accumulate_entries<K>(en[u], w[u], acc_q);
                        __syncwarp();   // the next edge of this slot may hit the same column from another lane
                    }
                }
            } else {
                for (int j = 0; j < nj; ++j) {
                    const int2 e = read_window(sv, cw, s0 + j);
                    if (s0 + j < sv.cnt) accumulate_any_k(cval, csel, (size_t)e.x * k, k, t, __int_as_float(e.y), acc_q);
                    __syncwarp();
                }
            }
        }
        if (has_nxt && !next_fetched) {               // an item without edges
            const WindowMap wn = make_window_map(d_nxt, sh_nxt, lane);
            fetch_window(wn, idx, val, 0, n_edges, pc, pw);
        }
        if (has_nxt) wm = make_window_map(d_nxt, sh_nxt, lane);
        __syncwarp();

        // ---- epilogue: write the rows of this item, re-zero the copies ---------------------------------
        if (it_cur.shared) {
            // lane l owns columns [8 l, 8 l + 8): word rows 2 l and 2 l + 1 of every copy, copies visited
            // in a lane-skewed order (4 lanes per bank group: the minimum for 16-byte accesses)
            const int r = __shfl_sync(kFullMask, d_cur.r, 0);
            float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
#pragma unroll
            for (int qq = 0; qq < kSS; ++qq) {
                const int w0 = (2 * lane) * 8 + ((qq + lane) & (kSS - 1));       // float4 index
                const float4 x = acc4[w0], y = acc4[w0 + 8];
                acc4[w0] = make_float4(0.f, 0.f, 0.f, 0.f);
                acc4[w0 + 8] = make_float4(0.f, 0.f, 0.f, 0.f);
                a.x += x.x; a.y += x.y; a.z += x.z; a.w += x.w;
                b.x += y.x; b.y += y.y; b.z += y.z; b.w += y.w;
            }
            float o[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
            if (row_div != nullptr) {
                const float dv = __ldg(row_div + r);
#pragma unroll
                for (int i = 0; i < 8; ++i) o[i] = div_guarded(o[i], dv);
            }
            float *orow = out + (size_t)r * ld_out;
            if (vec_out) {
                st_stream_f32x8(orow + 8 * lane, o);
            } else {
#pragma unroll
                for (int i = 0; i < 8; ++i)
                    if (8 * lane + i < dim) orow[8 * lane + i] = o[i];
            }
        } else {
            // lane l reads 8 consecutive columns [32 n + 8 te, +8) of copy qe = row of slot qe, with qe = l % 8 and
            // te = l / 8: a 16-byte shared-memory access is served a quarter-warp (8 consecutive lanes) at a time
            // and copy qe lives in banks [4 qe, 4 qe + 4), so the 8 lanes of a quarter must read 8 DIFFERENT copies
            // (the accumulation's own mapping, 4 lanes per copy, made every access here a 4-way conflict:
            // ncu 16 wavefronts per instruction instead of 4, 48 of the 137 shared-memory wavefronts per row
            // of the Yelp shape)
            const int qe = lane & (kSS - 1), te = lane >> 3;
            const int r = __shfl_sync(kFullMask, d_cur.r, qe);
            float dv = 1.f;
            const bool has_div = row_div != nullptr && r >= 0;
            if (has_div) dv = __ldg(row_div + r);
            float *orow = out + (size_t)(r >= 0 ? r : 0) * ld_out;
#pragma unroll
            for (int n = 0; n < kAccDim / 32; ++n) {
                const int w0 = (8 * n + 2 * te) * 8 + qe;                         // float4 index
                const float4 x = acc4[w0], y = acc4[w0 + 8];
                acc4[w0] = make_float4(0.f, 0.f, 0.f, 0.f);
                acc4[w0 + 8] = make_float4(0.f, 0.f, 0.f, 0.f);
                float o[8] = {x.x, x.y, x.z, x.w, y.x, y.y, y.z, y.w};
                if (has_div) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) o[i] = div_guarded(o[i], dv);
                }
                if (r >= 0) {
                    const int c0 = 32 * n + 8 * te;
                    if (vec_out) {
                        st_stream_f32x8(orow + c0, o);
                    } else {
#pragma unroll
                        for (int i = 0; i < 8; ++i)
                            if (c0 + i < dim) orow[c0 + i] = o[i];
                    }
                }
            }
        }
        __syncwarp();

        if (!has_nxt) break;
        it_cur.shared = sh_nxt;
        d_cur = d_nxt;
        if (has_nxt2) {
            d_nxt = d_nxt2;
            sh_nxt = sh_nxt2;
            has_nxt2 = false;
        } else if (!exhausted) {
            const int i = total_warps + grab_item(ticket, lane);
            has_nxt = i < pv.n_items;
            exhausted = !has_nxt;
            if (has_nxt) {
                const Item it = decode_item(pv, i);
                d_nxt = load_desc(pv, it, lane);
                sh_nxt = it.shared;
            }
        } else {
            has_nxt = false;
        }
    }
    leave_scheduler(ticket, lane, total_warps);
}

// scheduler ticket slot of the next launch (launches that overlap on one plan must not share a slot)
static int next_ticket_slot()
{
    static std::atomic<unsigned> counter{0};
    return (int)(counter.fetch_add(1) % kPlanTicketSlots);
}

template <int K, int MINB>
static cudaError_t launch_fwd_slots_b(const int *plan, const int *idx, const float *val, const float *cval,
                                      const uint8_t *csel, float *out, int n_edges, int dim, int ld_out, int k,
                                      const float *row_div, cudaStream_t stream)
{
    const size_t smem = (size_t)kFsWarps * kSlotWarpBytes;
    static LaunchConfig cache[kMaxCachedDevices];
    int dev = 0;
    cudaError_t err = cudaGetDevice(&dev);
    if (err != cudaSuccess) return err;
    LaunchConfig uncached = {false, 1, kNumSMsB200};
    LaunchConfig &cfg = (dev >= 0 && dev < kMaxCachedDevices) ? cache[dev] : uncached;
    if (!cfg.configured) {
        err = cudaFuncSetAttribute(spgemm_fwd_slots_kernel<K, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (err != cudaSuccess) return err;
        int blocks = 0;
        err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks, spgemm_fwd_slots_kernel<K, MINB>, kFsThreads, smem);
        if (err != cudaSuccess) return err;
        cfg.blocks_per_sm = blocks < 1 ? 1 : (blocks > MINB ? MINB : blocks);
        cfg.sms = device_sm_count();
        cfg.configured = true;
    }
    spgemm_fwd_slots_kernel<K, MINB><<<cfg.sms * cfg.blocks_per_sm, kFsThreads, smem, stream>>>(plan, idx, val, cval, csel,
                                                                                                out, n_edges, dim, ld_out, k, row_div, next_ticket_slot());
    return cudaGetLastError();
}

// Two CTAs of 8 warps per SM (128 registers, 4 steps = 32 edges gathered ahead): measured faster than three
// (80 registers, spills, and the 222 KB of shared memory leave the gathers almost no L1: Reddit shape 2.85 ms
// against 2.62 ms, ogbn-products shape 4.2 ms against 2.6 ms; profiles/r02_slots_lab.txt).
template <int K>
static cudaError_t launch_fwd_slots(const int *plan, const int *idx, const float *val, const float *cval,
                                    const uint8_t *csel, float *out, int n_edges, int dim, int ld_out, int k,
                                    const float *row_div, cudaStream_t stream)
{
    return launch_fwd_slots_b<K, 2>(plan, idx, val, cval, csel, out, n_edges, dim, ld_out, k, row_div, stream);
}

}  // namespace maxk

using namespace maxk;

// out rows may be strided (ld_out floats apart): the wide-feature path (wide.cu) writes 256-column blocks of a
// [n_rows, D > 256] matrix in place.
int maxk_forward_strided(const void *plan, const int32_t *indices, const float *values, const float *cbsr_val,
                         const uint8_t *cbsr_sel, float *out, int64_t ld_out, int64_t n_rows, int64_t n_edges, int dim,
                         int k, const float *row_div, cudaStream_t stream)
{
    if (dim < 1 || dim > kAccDim) return MAXK_ERR_BAD_DIM;
    if (k < 1 || k > kAccDim) return MAXK_ERR_BAD_K;
    if (n_rows < 0 || n_edges < 0 || n_rows > INT32_MAX - 64 || n_edges > INT32_MAX || ld_out < dim || ld_out > INT32_MAX)
        return MAXK_ERR_SIZE;
    if (n_rows == 0) return MAXK_OK;
    if (!plan || !out) return MAXK_ERR_NULL;
    if (n_edges > 0 && (!indices || !values || !cbsr_val || !cbsr_sel)) return MAXK_ERR_NULL;
    if (((uintptr_t)plan) & 15) return MAXK_ERR_ALIGN;
    if (dim == kAccDim && ((uintptr_t)out & 31)) return MAXK_ERR_ALIGN;
    const bool fast = (k == 8 || k == 16 || k == 32 || k == 64 || k == 96 || k == 128) &&
                      !(((uintptr_t)cbsr_val & 31) | ((uintptr_t)cbsr_sel & 7));
    const int *p = reinterpret_cast<const int *>(plan);
    const int ld = (int)ld_out;
    cudaError_t err;
    switch (fast ? k : 0) {
        case 8: err = launch_fwd_slots<8>(p, indices, values, cbsr_val, cbsr_sel, out, (int)n_edges, dim, ld, k, row_div, stream); break;
        case 16: err = launch_fwd_slots<16>(p, indices, values, cbsr_val, cbsr_sel, out, (int)n_edges, dim, ld, k, row_div, stream); break;
        case 32: err = launch_fwd_slots<32>(p, indices, values, cbsr_val, cbsr_sel, out, (int)n_edges, dim, ld, k, row_div, stream); break;
        case 64: err = launch_fwd_slots<64>(p, indices, values, cbsr_val, cbsr_sel, out, (int)n_edges, dim, ld, k, row_div, stream); break;
        case 96: err = launch_fwd_slots<96>(p, indices, values, cbsr_val, cbsr_sel, out, (int)n_edges, dim, ld, k, row_div, stream); break;
        case 128: err = launch_fwd_slots<128>(p, indices, values, cbsr_val, cbsr_sel, out, (int)n_edges, dim, ld, k, row_div, stream); break;
        default: err = launch_fwd_slots<0>(p, indices, values, cbsr_val, cbsr_sel, out, (int)n_edges, dim, ld, k, row_div, stream); break;
    }
    return status_from_cuda(err);
}

extern "C" int maxk_spgemm_forward_planned(const void *plan, const int32_t *indices, const float *values,
                                           const float *cbsr_val, const uint8_t *cbsr_sel, float *out, int64_t n_rows,
                                           int64_t n_edges, int dim, int k, const float *row_div,
                                           maxk_stream_t stream_)
{
    return maxk_forward_strided(plan, indices, values, cbsr_val, cbsr_sel, out, dim, n_rows, n_edges, dim, k, row_div,
                                (cudaStream_t)stream_);
}


/* Scratch of the un-planned entry points: the forward builds its row plan here on every call, the backward keeps
 * its scheduler counters and long-row list here. */
static size_t align16(size_t v) { return (v + 15) / 16 * 16; }
extern "C" size_t maxk_plan_bytes(int64_t n_rows);
extern "C" size_t maxk_plan_workspace_bytes(int64_t n_rows);
extern "C" int maxk_plan_build(const int32_t *row_begin, const int32_t *row_end, int64_t n_rows, void *plan,
                               size_t plan_bytes, void *workspace, size_t workspace_bytes, maxk_stream_t stream);

extern "C" size_t maxk_spgemm_workspace_bytes(int64_t n_rows)
{
    if (n_rows < 0) n_rows = 0;
    const size_t bwd = sizeof(SchedWorkspace) + sizeof(int) * (size_t)n_rows + 16;
    const size_t fwd = align16(maxk_plan_bytes(n_rows)) + align16(maxk_plan_workspace_bytes(n_rows));
    return align16(bwd > fwd ? bwd : fwd);
}

extern "C" int maxk_spgemm_forward(const int32_t *row_begin, const int32_t *row_end, const int32_t *indices,
                                   const float *values, const float *cbsr_val, const uint8_t *cbsr_sel, float *out,
                                   int64_t n_rows, int64_t n_edges, int dim, int k, const float *row_div,
                                   void *workspace, size_t workspace_bytes, maxk_stream_t stream)
{
    if (dim < 1 || dim > kAccDim) return MAXK_ERR_BAD_DIM;
    if (k < 1 || k > kAccDim) return MAXK_ERR_BAD_K;
    if (n_rows < 0 || n_edges < 0 || n_rows > INT32_MAX - 64 || n_edges > INT32_MAX) return MAXK_ERR_SIZE;
    if (n_rows == 0) return MAXK_OK;
    if (!row_begin || !row_end || !out || !workspace) return MAXK_ERR_NULL;
    if (workspace_bytes < maxk_spgemm_workspace_bytes(n_rows)) return MAXK_ERR_WORKSPACE;
    if ((uintptr_t)workspace & 15) return MAXK_ERR_ALIGN;
    const size_t pb = maxk_plan_bytes(n_rows), wb = maxk_plan_workspace_bytes(n_rows);
    unsigned char *base = reinterpret_cast<unsigned char *>(workspace);
    const int st = maxk_plan_build(row_begin, row_end, n_rows, base, pb, base + align16(pb), wb, stream);
    if (st != MAXK_OK) return st;
    return maxk_spgemm_forward_planned(base, indices, values, cbsr_val, cbsr_sel, out, n_rows, n_edges, dim, k, row_div,
                                       stream);
}
