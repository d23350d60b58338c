// slots.cuh -- the slot-parallel execution scheme of the forward SpGEMM (spgemm_fwd.cu).  (The backward SSpMM keeps
// 8 lanes per edge: one vector reduction must cover a destination row's full 128-byte line, see sspmm_bwd.cu.)
//
// A warp is split into kSS = 8 edge SLOTS of kSL = 4 lanes.  A slot processes one edge per step; the
// 4 lanes of a slot share the k entries of that edge's CBSR row (k/4 entries per lane).  Every slot has
// a private copy of the 256 columns (forward: accumulator, backward: staged gradient row) that lives
// only in the 4 shared-memory banks [4q, 4q+4):
//     word(column c, slot q) = (c >> 2) * 32 + 4q + (c & 3)
// so lanes of different slots can never collide, and inside a slot a bank conflict needs two of its
// 4 lanes to hold columns with equal c mod 4 in the same instruction.  The top-k kernel emits rows in
// MAXK_ORDER_BANKED (residue classes mod 4, largest class first, columns ascending inside a class; lane
// t takes k/4 consecutive entries), which keeps the conflicts of the 8 slots aligned on the same few
// instructions: simulated (tools/bank_conflict_sim.py) and measured wavefronts per edge are in DESIGN.md.
//
// Work items come from the row PLAN (plan.cu): either 8 different rows, one per slot ("separate":
// the epilogue writes 8 rows from the 8 copies, nothing is summed across copies and short rows fill
// all 32 lanes), or one row whose edges are dealt round-robin to the 8 slots ("shared": long rows and
// the final wave, the epilogue sums the 8 copies).  All slots of a warp advance in lockstep; the CSR
// (index, value) pairs of the next 16 steps of every slot are fetched with coalesced loads, one
// half-warp per slot, and handed to the slots through a small shared-memory window.
#pragma once
#include "maxk_common.cuh"

namespace maxk {

constexpr int kSL = 4;                      // lanes per slot == banks per copy
constexpr int kSS = 32 / kSL;               // slots per warp
constexpr int kSW = 16;                     // steps per CSR window
constexpr int kCwStride = 2 * kSW + 1;      // padded ring row (2 windows): 8 slots read 8 different bank pairs
constexpr int kCwEntries = kSS * kCwStride; // int2 entries per warp (>= kSS * kSW for the shared mode)
constexpr int kSlotCopyWords = kSS * kAccDim;                                   // 2048 floats per warp
constexpr size_t kSlotWarpBytes = kSlotCopyWords * sizeof(float) + kCwEntries * sizeof(int2);

constexpr int kPlanMagic = 0x4d41584b;      // "MAXK"
constexpr int kPlanHeaderInts = 16;
constexpr int kPlanLongDeg = 4096;          // rows with at least this many edges run in shared mode
constexpr int kPlanTailDeg = 128;           // final-wave rows at least this long run in shared mode
constexpr int kPlanTicketSlots = 32;        // (next item, finished warps) pairs behind the plan arrays: the dynamic
                                            // scheduler state of up to 32 launches in flight on one plan

// Row plan (device memory, int32): header, then three arrays of n_pad ints (row id, first edge, last edge
// + 1) in PLAN ORDER = rows sorted by degree bucket, longest first, stable inside a bucket.
//   items [0, nA)            shared   : plan position i                       (deg >= kPlanLongDeg)
//   items [nA, nA+nB)        separate : positions posB + 8 g .. (+8, clipped at posC)
//   items [.., +nC)          shared   : positions posC + j                    (the evening-out rows of a regular graph)
//   items [.., +nD)          separate : positions posD + 8 g .. (+8, clipped at n_rows)
struct PlanView {
    int n_rows, nA, nB, nC, nD, posC, posD, n_items;
    const int *p_row, *p_beg, *p_end;
    int *tickets;
};

__host__ __device__ inline int64_t plan_pad_rows(int64_t n_rows) { return (n_rows + 31) / 32 * 32 + 32; }

// Work items are handed out in plan order (heaviest first) by one atomic ticket counter per launch; the last warp
// to leave puts the counter pair back to zero, so a launch needs no memset and no workspace of its own.
__device__ __forceinline__ int grab_item(int *ticket, int lane)
{
    int v = 0;
    if (lane == 0) v = atomicAdd(ticket, 1);
    return __shfl_sync(kFullMask, v, 0);
}
__device__ __forceinline__ void leave_scheduler(int *ticket, int lane, int total_warps)
{
    if (lane == 0) {
        __threadfence();
        if (atomicAdd(ticket + 1, 1) == total_warps - 1) {
            ticket[0] = 0;
            ticket[1] = 0;
        }
    }
}

__device__ __forceinline__ PlanView load_plan(const int *plan)
{
    PlanView v;
    v.n_rows = __ldg(plan + 1);
    v.nA = __ldg(plan + 2);
    v.nB = __ldg(plan + 3);
    v.nC = __ldg(plan + 4);
    v.nD = __ldg(plan + 5);
    v.posC = __ldg(plan + 6);
    v.posD = __ldg(plan + 7);
    v.n_items = __ldg(plan + 8);
    const int64_t n_pad = plan_pad_rows(v.n_rows);
    v.p_row = plan + kPlanHeaderInts;
    v.p_beg = v.p_row + n_pad;
    v.p_end = v.p_beg + n_pad;
    v.tickets = const_cast<int *>(v.p_end + n_pad);
    return v;
}

struct Item {
    int shared;   // 1: one row dealt to the 8 slots; 0: up to 8 rows, one per slot
    int pos;      // first plan position
    int cnt;      // rows of the item (1 for shared)
};

__device__ __forceinline__ Item decode_item(const PlanView &p, int i)
{
    Item it;
    if (i < p.nA) {
        it.shared = 1; it.pos = i; it.cnt = 1;
    } else if (i < p.nA + p.nB) {
        it.shared = 0; it.pos = p.nA + kSS * (i - p.nA); it.cnt = min(kSS, p.posC - it.pos);
    } else if (i < p.nA + p.nB + p.nC) {
        it.shared = 1; it.pos = p.posC + (i - p.nA - p.nB); it.cnt = 1;
    } else {
        it.shared = 0; it.pos = p.posD + kSS * (i - p.nA - p.nB - p.nC); it.cnt = min(kSS, p.n_rows - it.pos);
    }
    return it;
}

// (row, begin, end) of up to 8 plan positions, held by lanes 0..7 (other lanes: row = -1, empty range)
struct Desc {
    int r, b, e;
};
__device__ __forceinline__ Desc load_desc(const PlanView &p, const Item &it, int lane)
{
    Desc d;
    d.r = -1; d.b = 0; d.e = 0;
    if (lane < it.cnt) {
        d.r = __ldg(p.p_row + it.pos + lane);
        d.b = __ldg(p.p_beg + it.pos + lane);
        d.e = __ldg(p.p_end + it.pos + lane);
    }
    return d;
}

// CSR windows of one item.  Window w of a slot holds 16 consecutive (index, value) pairs; fill instruction i
// (4 per window) fetches the windows of slots 2i and 2i+1, one half-warp each, and parks them in a two-window
// ring per slot in shared memory.
//   * short items: window w = the slot's edges [b + 16 w, b + 16 w + 16); neighbouring rows of a regular
//     low-degree graph are contiguous in the CSR arrays, so the 8 windows of a warp share their sectors;
//   * long items ("aligned", >= kAlignedSteps steps): windows start at multiples of 16 entries (64 bytes), the
//     DRAM access granularity -- un-aligned 64-byte windows made every 64-byte chunk of the CSR arrays cross
//     the DRAM bus twice (ncu, Reddit shape: 1.63 GB read against 1.0 GB of operands).  A slot then reads its
//     step s at ring entry (b % 16 + s) % 32 and needs windows j and j+1 parked during block j (lead = 1);
//   * shared items: one row, 128 consecutive edges per 16 steps, flat (slot q, step s) -> entry 8 s + q.
constexpr int kAlignedSteps = 64;
constexpr int kRing = 2 * kSW;              // ring entries per slot
struct WindowMap {
    int base[kSS / 2];   // edge position fetched for window 0
    int wi0, wis;        // window entry written by fill instruction i: wi0 + i * wis (+ ring offset of the window)
    int shared;          // edge positions per window: 16 (separate) / 128 (shared); ring offset of odd windows: 16 / 0
    int lead;            // windows that must be parked ahead of the block being processed (1 aligned, else 0)
    int n_win;           // windows of the item
};
__device__ __forceinline__ WindowMap make_window_map(const Desc &d, int shared, int lane)
{
    WindowMap m;
    const int half = lane >> 4, l16 = lane & 15;
    if (shared) {
        const int b0 = __shfl_sync(kFullMask, d.b, 0), e0 = __shfl_sync(kFullMask, d.e, 0);
#pragma unroll
        for (int i = 0; i < kSS / 2; ++i) m.base[i] = b0 + 32 * i + lane;
        m.wi0 = lane;
        m.wis = 32;
        m.shared = 1;
        m.lead = 0;
        m.n_win = ((e0 - b0 + kSS - 1) / kSS + kSW - 1) / kSW;
    } else {
        const int steps = __reduce_max_sync(kFullMask, d.e - d.b);
        const int aligned = steps >= kAlignedSteps;
#pragma unroll
        for (int i = 0; i < kSS / 2; ++i) {
            const int wb = __shfl_sync(kFullMask, d.b, 2 * i + half);
            m.base[i] = (aligned ? (wb & ~(kSW - 1)) : wb) + l16;
        }
        m.wi0 = half * kCwStride + l16;
        m.wis = 2 * kCwStride;
        m.shared = 0;
        m.lead = aligned;
        m.n_win = (steps + (kSW - 1) * (1 + aligned)) / kSW;
    }
    return m;
}

// Window loads bypass L1.  L2 eviction hint (compile-time, measured on B200: 0 = none, 1 = evict_normal,
// 2 = evict_last, 3 = evict_first): the DRAM bytes of the Reddit-shape forward do not depend on it (1.33-1.35 GB
// read), evict_normal / evict_last are ~1 % faster on the Yelp and products shapes (0.566 vs 0.573 ms, 2.52 vs
// 2.54 ms) where the CBSR competes with the streams for L2.
#ifndef MAXK_WINDOW_POLICY
#define MAXK_WINDOW_POLICY 1
#endif
__device__ __forceinline__ uint64_t window_policy()
{
    uint64_t p = 0;
#if MAXK_WINDOW_POLICY == 1
    asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(p));
#elif MAXK_WINDOW_POLICY == 2
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
#elif MAXK_WINDOW_POLICY == 3
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
#endif
    return p;
}
__device__ __forceinline__ int ld_window_i32(const int *p)
{
    int v;
#if MAXK_WINDOW_POLICY == 0
    asm volatile("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(v) : "l"(p));
#else
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.s32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(window_policy()));
#endif
    return v;
}
__device__ __forceinline__ float ld_window_f32(const float *p)
{
    float v;
#if MAXK_WINDOW_POLICY == 0
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
#else
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(v) : "l"(p), "l"(window_policy()));
#endif
    return v;
}

// fetch window w (register-resident until it is parked); positions past the edge arrays are skipped, positions
// past the slot's own row are fetched and never read
__device__ __forceinline__ void fetch_window(const WindowMap &m, const int *__restrict__ idx,
                                             const float *__restrict__ val, int w, int n_edges, int (&pc)[kSS / 2],
                                             float (&pw)[kSS / 2])
{
#pragma unroll
    for (int i = 0; i < kSS / 2; ++i) {
        const int a = m.base[i] + w * (m.shared ? kSS * kSW : kSW);
        pc[i] = 0;
        pw[i] = 0.f;
        if (a < n_edges) {
            pc[i] = ld_window_i32(idx + a);
            pw[i] = ld_window_f32(val + a);
        }
    }
}

__device__ __forceinline__ void park_window(const WindowMap &m, int2 *cw, int w, const int (&pc)[kSS / 2],
                                            const float (&pw)[kSS / 2])
{
    const int off = m.wi0 + (m.shared ? 0 : (w & 1) * kSW);
#pragma unroll
    for (int i = 0; i < kSS / 2; ++i) cw[off + i * m.wis] = make_int2(pc[i], __float_as_int(pw[i]));
}

// per-lane view of its slot inside an item
struct SlotView {
    int cnt;        // steps this slot works
    int steps;      // steps of the item (max over slots)
    int rd_base;    // window entry of ring position 0
    int rd_stride;  // window entries per ring position
    int rd_mask;    // ring positions - 1
    int rd_off;     // ring position of step 0
};
__device__ __forceinline__ SlotView make_slot_view(const Desc &d, int shared, int lane)
{
    SlotView s;
    const int q = lane / kSL;
    if (shared) {
        const int len = __shfl_sync(kFullMask, d.e - d.b, 0);
        s.cnt = len > q ? (len - q + kSS - 1) / kSS : 0;
        s.steps = (len + kSS - 1) / kSS;
        s.rd_base = q;
        s.rd_stride = kSS;
        s.rd_mask = kSW - 1;
        s.rd_off = 0;
    } else {
        const int len = d.e - d.b;                       // lanes >= 8 hold 0
        s.cnt = __shfl_sync(kFullMask, len, q);
        s.steps = __reduce_max_sync(kFullMask, len);
        s.rd_base = q * kCwStride;
        s.rd_stride = 1;
        s.rd_mask = kRing - 1;
        s.rd_off = s.steps >= kAlignedSteps ? (__shfl_sync(kFullMask, d.b, q) & (kSW - 1)) : 0;
    }
    return s;
}
__device__ __forceinline__ int2 read_window(const SlotView &s, const int2 *cw, int step)
{
    return cw[s.rd_base + ((s.rd_off + step) & s.rd_mask) * s.rd_stride];
}

// word offset of column c inside a slot copy (add 4 * slot)
__device__ __forceinline__ int slot_word(int c) { return ((c & 0xfc) << 3) | (c & 3); }

// 32-byte / 8-byte loads of re-used operands (kept in L2)
__device__ __forceinline__ void ld_keep_f32x8(const float *p, float *v, uint64_t pol)
{
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8], %9;"
                 : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7])
                 : "l"(p), "l"(pol));
}
__device__ __forceinline__ uint2 ld_keep_u32x2(const void *p, uint64_t pol)
{
    uint2 v;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v2.u32 {%0,%1}, [%2], %3;" : "=r"(v.x), "=r"(v.y) : "l"(p), "l"(pol));
    return v;
}

// Entries of one edge owned by one lane (t = lane & 3).  Memory position of the lane's entries inside
// the CBSR row:  k >= 32: 32-entry chunks, lane t owns entries [32 j + 8 t, +8) of chunk j (one 32-byte
// value load + one 8-byte selector load per chunk, every 128-byte line touched once per instruction);
// k = 16: [4 t, +4);  k = 8: [2 t, +2).
template <int K> struct SlotEntries {
    static_assert(K == 8 || K == 16 || (K % 32 == 0 && K >= 32 && K <= 128), "unsupported k");
    static constexpr int EPL = K / kSL;
    static constexpr int CH = K >= 32 ? K / 32 : 1;
    float v[EPL];
    uint32_t s[(EPL + 3) / 4];
    __device__ __forceinline__ int col(int i) const { return (s[i >> 2] >> (8 * (i & 3))) & 0xff; }
    // offset (in entries) of this lane's first entry of chunk j
    __device__ static __forceinline__ int first(int t, int j) { return K >= 32 ? 32 * j + 8 * t : EPL * t; }
    __device__ __forceinline__ void load_sel(const uint8_t *csel, size_t row_off, int t, uint64_t keep)
    {
        if constexpr (K >= 32) {
#pragma unroll
            for (int j = 0; j < CH; ++j) {
                const uint2 u = ld_keep_u32x2(csel + row_off + first(t, j), keep);
                s[2 * j] = u.x;
                s[2 * j + 1] = u.y;
            }
        } else if constexpr (K == 16) {
            s[0] = ld_keep_u32(csel + row_off + 4 * t, keep);
        } else {
            s[0] = ld_keep_u16(csel + row_off + 2 * t, keep);
        }
    }
    __device__ __forceinline__ void load_val(const float *cval, size_t row_off, int t, uint64_t keep)
    {
        if constexpr (K >= 32) {
#pragma unroll
            for (int j = 0; j < CH; ++j) ld_keep_f32x8(cval + row_off + first(t, j), v + 8 * j, keep);
        } else if constexpr (K == 16) {
            const float4 a = ld_keep_f32x4(cval + row_off + 4 * t, keep);
            v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
        } else {
            const float2 a = ld_keep_f32x2(cval + row_off + 2 * t, keep);
            v[0] = a.x; v[1] = a.y;
        }
    }
};

}  // namespace maxk
