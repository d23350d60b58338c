// topk.cu -- MaxK row-wise top-k -> CBSR, plus the dense helpers of the MaxK nonlinearity (sm_100a).
//
// Replaces torch.topk(x, k, dim=1) + .to(uint8) on the reference's training path
// (maxk_spgemm_function.py:53-57, model_integrated_v3.py:28-37) and the uint8 `topk` kernel
// (kernels/maxk_kernel.cu:23-96), which quantises to 8 bits, emits the FIRST k elements above a
// pivot rather than the top k and is hard-wired to k = 32.
//
// One warp per row, the row lives in registers (8 values per lane).  Exact selection:
//   1. keys  = order-preserving uint32 image of the floats (NaN largest, -0 == +0);
//   2. a threshold search in KEY space: the bracket starts at the minimum over lanes of each lane's
//      ceil(k/32)-th largest key (at least k keys lie at or above it), pivots come from a secant step on
//      the counts (a plain bisection every fourth step), count(key >= pivot) is one warp REDUX per step;
//      it stops as soon as exactly k keys lie at or above the pivot (5.3 counts per row on U[0,1) data,
//      6.5 on N(0,1); at most ~128 when ties straddle rank k);
//   3. every key > T is selected; of the keys == T the lowest columns are taken until k;
//   4. the output position of every selected entry is computed from 8 warp ballots with
//      popcounts: column order (MAXK_ORDER_COLUMN_ASC), bank-residue-major order (MAXK_ORDER_BANKED,
//      what the SpGEMM/SSpMM kernels are conflict-minimal on), or rank-sorted by (value desc, column
//      asc) through shared memory (MAXK_ORDER_VALUE_DESC).
// Two kernels: topk_banked_kernel<K, PLAIN> for the layer's hot configuration (dim 256, banked order,
// k in {8, 16, 32, 64}; see its header below) and the general topk_cbsr_kernel for everything else.
// The same pass can write the dense masked row (the MaxK nonlinearity output), so the
// reference's topk + zeros_like + scatter_ + multiply (4 dense passes) is one read + one write.
#include "maxk_common.cuh"

namespace maxk {

constexpr int kTopkThreads = 256;
constexpr int kTopkWarps = kTopkThreads / 32;
constexpr unsigned kFullT = 0xffffffffu;

// Column owned by (lane, slot): slots 0-3 -> 4*lane+slot, slots 4-7 -> 128+4*lane+(slot-4).
__device__ __forceinline__ int col_of(int lane, int slot) { return (slot < 4 ? 0 : 128) + 4 * lane + (slot & 3); }

// Order-preserving key via one FADD: v + 0.0f turns -0 into +0 and every NaN into the canonical
// positive quiet NaN 0x7fffffff (largest key), so no compare/select is needed per element.
__device__ __forceinline__ uint32_t fast_key(float v)
{
    const uint32_t b = __float_as_uint(__fadd_rn(v, 0.0f));
    return b ^ ((uint32_t)((int32_t)b >> 31) | 0x80000000u);
}

// MAXK_ORDER_BANKED: sorted rank p (classes mod 4 by size, columns ascending) -> position in the CBSR row.
// Lane t of a slot (slots.cuh) takes the k/4 consecutive ranks [t k/4, (t+1) k/4) and processes the i-th of
// them in its i-th instruction; for k >= 32 its entries are stored as 8-entry runs at 32 j + 8 t (one
// 32-byte load per 32-entry chunk j), so rank p = t k/4 + 8 j + e lives at 32 j + 8 t + e.
__host__ __device__ inline int banked_mem_pos(int p, int k)
{
    if (k < 32) return p;
    const int epl = k >> 2, t = p / epl, i = p - t * epl;
    return 32 * (i >> 3) + 8 * t + (i & 7);
}

// c_recip31[d] = floor(2^31 / d), d in [1, 256]; statically initialised so that no host-side copy is
// needed (the first call may happen inside a CUDA-graph capture)
__constant__ uint32_t c_recip31[kAccDim + 1] = {
#include "recip31_table.inc"
};

// Warp-wide count of keys >= t.  The search is bound by the ALU pipe (ncu: 75 % of its peak, one warp
// instruction per two cycles; the FMA pipe 17 %), so the count is split over both pipes: the compare is an
// ALU-pipe ISETP, the increment a predicated IMAD c = one * one + c on the FMA pipe.  `one` is a kernel
// argument holding 1 -- with a literal the compiler folds the multiply into an ALU-pipe add again
// (tools/count_bench.cu: 30 cycles per count per scheduler against 39 for what `c += key >= t` compiles to
// and 43 for a borrow chain of IADD3 pairs).
__device__ __forceinline__ int count_ge(const uint32_t (&key)[8], uint32_t t, int one)
{
    int c = 0;
    asm("{\n\t.reg .pred p;\n\t"
        "setp.ge.u32 p, %1, %9;\n\t@p mad.lo.s32 %0, %10, %10, %0;\n\t"
        "setp.ge.u32 p, %2, %9;\n\t@p mad.lo.s32 %0, %10, %10, %0;\n\t"
        "setp.ge.u32 p, %3, %9;\n\t@p mad.lo.s32 %0, %10, %10, %0;\n\t"
        "setp.ge.u32 p, %4, %9;\n\t@p mad.lo.s32 %0, %10, %10, %0;\n\t"
        "setp.ge.u32 p, %5, %9;\n\t@p mad.lo.s32 %0, %10, %10, %0;\n\t"
        "setp.ge.u32 p, %6, %9;\n\t@p mad.lo.s32 %0, %10, %10, %0;\n\t"
        "setp.ge.u32 p, %7, %9;\n\t@p mad.lo.s32 %0, %10, %10, %0;\n\t"
        "setp.ge.u32 p, %8, %9;\n\t@p mad.lo.s32 %0, %10, %10, %0;\n\t}"
        : "+r"(c)
        : "r"(key[0]), "r"(key[1]), "r"(key[2]), "r"(key[3]), "r"(key[4]), "r"(key[5]), "r"(key[6]), "r"(key[7]), "r"(t),
          "r"(one));
    return __reduce_add_sync(kFullT, c);
}

// Threshold T with count(key > T) <= k <= count(key >= T), given a bracket: count(>= kmin) >= k and kmax
// the largest key of the row.  Invariant: count(>= lo) = c_lo >= k, count(>= hi) = c_hi < k.  Pivots come
// from linear interpolation of the count (few steps on smooth data) with a plain bisection every fourth
// step (guaranteed progress on anything); stops early when exactly k keys are at or above the pivot
// (`exact`: count(>= returned T) == k, no tie handling needed).  (p2, c2) is an optional second known point,
// count(>= p2) = c2 (pass p2 = 0 for none): it replaces whichever end of the bracket it tightens.
__device__ __forceinline__ uint32_t find_threshold(const uint32_t (&key)[8], int k, uint32_t kmin, uint32_t kmax,
                                                   bool &exact, int one, uint32_t p2 = 0u, int c2 = 0)
{
    exact = false;
    if (kmax == 0xffffffffu && count_ge(key, kmax, one) >= k) return kmax;   // (NaN rows only) rank k lies inside the run of maximal keys
    // hi is exclusive: nothing is >= kmax + 1 (kmax = 0xffffffff keeps hi = kmax, whose count is < k here)
    uint32_t lo = kmin, hi = kmax == 0xffffffffu ? kmax : kmax + 1u;
    int c_lo = count_ge(key, lo, one);
    int c_hi = kmax == 0xffffffffu ? count_ge(key, kmax, one) : 0;
    if (p2 > lo && p2 < hi) {
        if (c2 >= k) { lo = p2; c_lo = c2; } else { hi = p2; c_hi = c2; }
    }
    // One count per step.  A pivot with exactly k keys at or above it collapses the bracket onto itself
    // (lo = hi = pivot), so the span test alone ends the loop; every fourth step is a plain bisection, which
    // bounds the search on anything (<= 4 * 32 steps) -- written as four copies of the step instead of a
    // step counter (selects, not branches: straight-line code around one count; the ALU pipe is the bound).
    uint32_t span = c_lo == k ? 0u : hi - lo;
    auto step = [&](bool bisect) {
        // secant step: off = span * (c_lo - k + 1/2) / (c_lo - c_hi) in integers; counts are <= 256
        // (num < 2 * d, so num * floor(2^31 / d) < 2^32 is the fraction in Q32)
        const uint32_t num = (uint32_t)(2 * (c_lo - k) + 1);
        const uint32_t sec = bisect ? 0u : min(max(__umulhi(span, num * c_recip31[c_lo - c_hi]), 1u), span - 1u);
        const uint32_t mid = lo + (bisect ? (span >> 1) : sec);
        const int c = count_ge(key, mid, one);
        const bool up = c >= k, dn = c <= k;     // both when c == k
        lo = up ? mid : lo;
        c_lo = up ? c : c_lo;
        hi = dn ? mid : hi;
        c_hi = dn ? c : c_hi;
        span = hi - lo;
    };
    for (;;) {
        if (span <= 1u) break;
        step(false);
        if (span <= 1u) break;
        step(false);
        if (span <= 1u) break;
        step(false);
        if (span <= 1u) break;
        step(true);
    }
    exact = c_lo == k;                           // count(>= lo) == k: no tie handling needed
    return lo;
}

// DIM256: dim == 256 (two 16-byte loads per lane, no padding logic).  ORDER: MAXK_ORDER_*.
template <bool DIM256, int ORDER>
__global__ void __launch_bounds__(kTopkThreads)
topk_cbsr_kernel(const float *__restrict__ x, int64_t n_rows, int dim, int k, int bank_mod,
                 float *__restrict__ out_val, uint8_t *__restrict__ out_sel, int32_t *__restrict__ out_i32,
                 int64_t *__restrict__ out_i64, float *__restrict__ masked, int one)
{
    __shared__ float s_val[kTopkWarps][kAccDim];
    __shared__ uint8_t s_col[kTopkWarps][kAccDim];
    __shared__ uint32_t s_key[ORDER == MAXK_ORDER_VALUE_DESC ? kTopkWarps : 1][kAccDim];
    const int lane = lane_id();
    const int warp = threadIdx.x >> 5;
    const int64_t warps_total = (int64_t)gridDim.x * kTopkWarps;
    const unsigned lt = (1u << lane) - 1u;

    for (int64_t r = (int64_t)blockIdx.x * kTopkWarps + warp; r < n_rows; r += warps_total) {
        const float *row = x + r * dim;
        float v[8];
        if (DIM256) {
            const float4 a = ld_stream_f32x4(row + 4 * lane);
            const float4 b = ld_stream_f32x4(row + 128 + 4 * lane);
            v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
            v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
        } else {
#pragma unroll
            for (int s = 0; s < 8; ++s) {
                const int c = col_of(lane, s);
                v[s] = c < dim ? row[c] : 0.f;
            }
        }
        uint32_t key[8];
        uint32_t kmax = 0u, kmin = 0xffffffffu;
#pragma unroll
        for (int s = 0; s < 8; ++s) {
            const bool real = DIM256 || col_of(lane, s) < dim;
            key[s] = real ? fast_key(v[s]) : 0u;    // pad: below every real key (real keys are >= 0x007fffff)
        }
        // Lower end of the search bracket.  Every lane has >= j keys at or above its own j-th largest
        // key, so lo = min over lanes of that key has count(>= lo) >= 32 j >= k for j = ceil(k / 32):
        // a far tighter start than the row minimum (U[0,1): ~90 keys above it instead of 256, and in
        // the same binade as the threshold, where keys are linear in the value) -- the secant search
        // then needs ~3 steps instead of ~7.
        uint32_t m1 = max(key[0], key[1]), m2 = min(key[0], key[1]);
#pragma unroll
        for (int s = 2; s < 8; ++s) {
            m2 = max(m2, min(m1, key[s]));
            m1 = max(m1, key[s]);
        }
        kmax = __reduce_max_sync(kFullT, m1);
        if (DIM256 && k <= 32) kmin = __reduce_min_sync(kFullT, m1);
        else if (DIM256 && k <= 64) kmin = __reduce_min_sync(kFullT, m2);
        else {
#pragma unroll
            for (int s = 0; s < 8; ++s)
                if (DIM256 || col_of(lane, s) < dim) kmin = min(kmin, key[s]);
            kmin = __reduce_min_sync(kFullT, kmin);
        }

        bool exact;
        const uint32_t T = find_threshold(key, k, kmin, kmax, exact, one);

        // ---- selection flags ---------------------------------------------------------------------
        bool selb[8];
        if (exact) {
#pragma unroll
            for (int s = 0; s < 8; ++s) selb[s] = key[s] >= T;
        } else {
            // ties straddle rank k: key > T always, key == T lowest column first, (half, lane, slot&3)
            unsigned b_eq[8];
            int gt_total = 0;
#pragma unroll
            for (int s = 0; s < 8; ++s) {
                gt_total += __popc(__ballot_sync(kFullT, key[s] > T));
                b_eq[s] = __ballot_sync(kFullT, key[s] == T);
            }
            const int need_eq = k - gt_total;
            int eq_lo_before = 0, eq_lo_total = 0, eq_hi_before = 0;
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                eq_lo_before += __popc(b_eq[u] & lt);
                eq_lo_total += __popc(b_eq[u]);
                eq_hi_before += __popc(b_eq[u + 4] & lt);
            }
            int rk = eq_lo_before;
#pragma unroll
            for (int s = 0; s < 4; ++s) {
                const bool is_eq = key[s] == T;
                selb[s] = (key[s] > T) || (is_eq && rk < need_eq);
                rk += is_eq;
            }
            rk = eq_lo_total + eq_hi_before;
#pragma unroll
            for (int s = 4; s < 8; ++s) {
                const bool is_eq = key[s] == T;
                selb[s] = (key[s] > T) || (is_eq && rk < need_eq);
                rk += is_eq;
            }
        }

        if (masked != nullptr) {
            float *mrow = masked + r * dim;
            if (DIM256) {
                st_stream_f32x4(mrow + 4 * lane, make_float4(selb[0] ? v[0] : 0.f, selb[1] ? v[1] : 0.f,
                                                            selb[2] ? v[2] : 0.f, selb[3] ? v[3] : 0.f));
                st_stream_f32x4(mrow + 128 + 4 * lane, make_float4(selb[4] ? v[4] : 0.f, selb[5] ? v[5] : 0.f,
                                                                  selb[6] ? v[6] : 0.f, selb[7] ? v[7] : 0.f));
            } else {
#pragma unroll
                for (int s = 0; s < 8; ++s) {
                    const int c = col_of(lane, s);
                    if (c < dim) mrow[c] = selb[s] ? v[s] : 0.f;
                }
            }
        }

        // ---- output positions from the selection ballots (popcounts only) ---------------------------
        unsigned bs[8];
#pragma unroll
        for (int s = 0; s < 8; ++s) bs[s] = __ballot_sync(kFullT, selb[s]);
        int pos[8];
        if (ORDER == MAXK_ORDER_BANKED && bank_mod >= 4) {
            // class of (lane, slot) = slot & 3 (= column mod 4); classes by (size desc, class asc), columns
            // ascending inside a class: all columns of the first half come before those of the second half
            int sz[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) sz[u] = __popc(bs[u]) + __popc(bs[u + 4]);
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                int base = 0;
#pragma unroll
                for (int v = 0; v < 4; ++v)
                    if (v != u && (sz[v] > sz[u] || (sz[v] == sz[u] && v < u))) base += sz[v];
                pos[u] = banked_mem_pos(base + __popc(bs[u] & lt), k);
                pos[u + 4] = banked_mem_pos(base + __popc(bs[u]) + __popc(bs[u + 4] & lt), k);
            }
        } else {
            // column order: (half, lane, slot&3)
            int lo_before = 0, lo_total = 0, hi_before = 0;
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                lo_before += __popc(bs[u] & lt);
                lo_total += __popc(bs[u]);
                hi_before += __popc(bs[u + 4] & lt);
            }
            int p = lo_before;
#pragma unroll
            for (int s = 0; s < 4; ++s) { pos[s] = p; p += selb[s] ? 1 : 0; }
            p = lo_total + hi_before;
#pragma unroll
            for (int s = 4; s < 8; ++s) { pos[s] = p; p += selb[s] ? 1 : 0; }
        }

        // ---- stage the k entries in shared memory (scattered 4-byte global stores measured slower),
        //      then coalesced stores; the value order rank-sorts the staged entries on the way out
#pragma unroll
        for (int s = 0; s < 8; ++s)
            if (selb[s]) {
                s_val[warp][pos[s]] = v[s];
                s_col[warp][pos[s]] = (uint8_t)col_of(lane, s);
                if (ORDER == MAXK_ORDER_VALUE_DESC) s_key[warp][pos[s]] = key[s];
            }
        __syncwarp();

        for (int i = lane; i < k; i += 32) {
            int dst = i;
            if (ORDER == MAXK_ORDER_VALUE_DESC) {
                // entries are staged in column order, so "earlier index" == "lower column"
                const uint32_t ki = s_key[warp][i];
                int rank = 0;
                for (int j = 0; j < k; ++j) {
                    const uint32_t kj = s_key[warp][j];
                    rank += (kj > ki) || (kj == ki && j < i);
                }
                dst = rank;
            }
            const int64_t o = r * k + dst;
            const int c = s_col[warp][i];
            out_val[o] = s_val[warp][i];
            if (out_sel) out_sel[o] = (uint8_t)c;
            if (out_i32) out_i32[o] = c;
            if (out_i64) out_i64[o] = c;
        }
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------------
// The layer's hot configuration: dim == 256, MAXK_ORDER_BANKED, k in {8, 16, 32, 64}.
//
// Same selection as the general kernel above, arranged for the fewest issued instructions (386 warp
// instructions per row for k = 32, ALU pipe 75 % busy; the general kernel issues ~565 and reaches 19 % of the
// HBM bandwidth): lane l owns the 8 CONSECUTIVE columns 8l .. 8l+7 (one 32-byte load per lane), so
// that a column's bank-residue class mod 8 is its register slot and the banked output position of an
// entry is a prefix over the slots' ballots instead of a lane-group exchange; (value, column) pairs are
// staged as one 8-byte shared-memory store; k is a template parameter.
// ---------------------------------------------------------------------------------------------
// Peer destinations of the row-sharded layer (sharded.py): the CBSR rows of this rank's slab are written
// straight into every rank's gathered [P * m, k] buffers (peer-mapped symmetric memory over NVLink) instead
// of a local store followed by two all_gather launches; the bytes on the wire are the same.
struct PeerOut {
    int n;                         // 0: plain local output
    float *val[MAXK_MAX_PEERS];    // already offset to this rank's first row
    uint8_t *sel[MAXK_MAX_PEERS];
    float *mc_val;                 // NVLS multicast mappings of the same buffers (nullable): ONE store reaches all ranks
    uint32_t *mc_sel;
};

__device__ __forceinline__ void multimem_st_b32(void *mc, uint32_t bits)
{
    asm volatile("multimem.st.relaxed.sys.global.f32 [%0], %1;" ::"l"(mc), "f"(__uint_as_float(bits)) : "memory");
}

// PLAIN: values + uint8 selectors to local memory only (what the layer's forward pass launches): the optional
// outputs (int32 / int64 indices, masked row, peer destinations) cost ~35 instructions per row in null checks
// and predicated address arithmetic on the pipe that bounds the kernel, so they are compiled out.
template <int K, bool PLAIN>
__global__ void __launch_bounds__(kTopkThreads)
topk_banked_kernel(const float *__restrict__ x, int64_t n_rows, float *__restrict__ out_val,
                   uint8_t *__restrict__ out_sel, int32_t *__restrict__ out_i32, int64_t *__restrict__ out_i64,
                   float *__restrict__ masked, int one, const PeerOut peers)
{
    __shared__ __align__(8) uint32_t s_ent[kTopkWarps][2 * K];   // (value bits, column id) pairs
    const int lane = lane_id();
    const int warp = threadIdx.x >> 5;
    const int64_t warps_total = (int64_t)gridDim.x * kTopkWarps;
    const unsigned lt = (1u << lane) - 1u;

    // The next row of the warp is loaded while the current one is searched: a warp that only loads between
    // rows keeps ~1/3 of its 1 KiB in flight on average, and 40 warps x 1/3 KiB per SM is far below the
    // ~44 KiB per SM that the HBM latency-bandwidth product asks for (measured: 1.7-2.0 TB/s without).
    int64_t r = (int64_t)blockIdx.x * kTopkWarps + warp;
    float nv[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (r < n_rows) ld_stream_f32x8(x + r * kAccDim + 8 * lane, nv);
    for (; r < n_rows; r += warps_total) {
        float v[8];
#pragma unroll
        for (int s = 0; s < 8; ++s) v[s] = nv[s];
        if (r + warps_total < n_rows) ld_stream_f32x8(x + (r + warps_total) * kAccDim + 8 * lane, nv);
        uint32_t key[8];
#pragma unroll
        for (int s = 0; s < 8; ++s) key[s] = fast_key(v[s]);
        // bracket: lo = min over lanes of the lane's ceil(K/32)-th largest key (see the general kernel)
        uint32_t m1 = max(key[0], key[1]), m2 = min(key[0], key[1]);
#pragma unroll
        for (int s = 2; s < 8; ++s) {
            if (K > 32 || K <= 16) m2 = max(m2, min(m1, key[s]));
            m1 = max(m1, key[s]);
        }
        const uint32_t kmax = __reduce_max_sync(kFullT, m1);
        uint32_t lane_lo = K > 32 ? m2 : m1;
        if (K > 64) {                           // no per-lane rank bound beyond the 2nd largest: start at the row minimum
            lane_lo = key[0];
#pragma unroll
            for (int s = 1; s < 8; ++s) lane_lo = min(lane_lo, key[s]);
        }
        const uint32_t kmin = __reduce_min_sync(kFullT, lane_lo);
        // small k: a second exact point for free.  Only lane maxima can exceed the largest second-largest
        // key M2 of any lane, so count(>= M2 + 1) is one ballot over the lane maxima; it cuts the long
        // tail above the threshold off the bracket (simulated: 4.7 -> 3.8 counts per row on U[0,1),
        // 8.4 -> 4.7 on N(0,1) for k = 8; no gain for k >= 32, where it is not computed).
        uint32_t p2 = 0u;
        int c2 = 0;
        if (K <= 16) {
            const uint32_t big2 = __reduce_max_sync(kFullT, m2);
            p2 = big2 + 1u;                       // wraps to 0 (= none) when big2 is the NaN key
            c2 = __popc(__ballot_sync(kFullT, m1 > big2));
        }
        bool exact;
        const uint32_t T = find_threshold(key, K, kmin, kmax, exact, one, p2, c2);

        bool selb[8];
        if (exact) {
#pragma unroll
            for (int s = 0; s < 8; ++s) selb[s] = key[s] >= T;
        } else {
            // ties straddle rank K: key > T always, key == T lowest column first = (lane, slot) order
            int gt_total = 0, rk = 0;
#pragma unroll
            for (int s = 0; s < 8; ++s) {
                gt_total += __popc(__ballot_sync(kFullT, key[s] > T));
                rk += __popc(__ballot_sync(kFullT, key[s] == T) & lt);
            }
            const int need_eq = K - gt_total;
#pragma unroll
            for (int s = 0; s < 8; ++s) {
                const bool is_eq = key[s] == T;
                selb[s] = (key[s] > T) || (is_eq && rk < need_eq);
                rk += is_eq ? 1 : 0;
            }
        }

        if (!PLAIN && masked != nullptr) {
            float mv[8];
#pragma unroll
            for (int s = 0; s < 8; ++s) mv[s] = selb[s] ? v[s] : 0.f;
            st_stream_f32x8(masked + r * kAccDim + 8 * lane, mv);
        }

        // ---- banked output positions: class = column mod M, columns ascending inside a class -------------
        unsigned bs[8];
#pragma unroll
        for (int s = 0; s < 8; ++s) bs[s] = __ballot_sync(kFullT, selb[s]);
        // class == slot & 3; inside a class: (lane, slot < 4 first); classes by (size desc, class asc)
        int pos[8];
        {
            int sz[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) sz[u] = __popc(bs[u]) + __popc(bs[u + 4]);
            // base[u] = total size of the classes ranked before u; class a < b comes first iff sz[a] >= sz[b]
            // (six warp-uniform compares)
            const bool f01 = sz[0] >= sz[1], f02 = sz[0] >= sz[2], f03 = sz[0] >= sz[3];
            const bool f12 = sz[1] >= sz[2], f13 = sz[1] >= sz[3], f23 = sz[2] >= sz[3];
            const int base[4] = {(f01 ? 0 : sz[1]) + (f02 ? 0 : sz[2]) + (f03 ? 0 : sz[3]),
                                 (f01 ? sz[0] : 0) + (f12 ? 0 : sz[2]) + (f13 ? 0 : sz[3]),
                                 (f02 ? sz[0] : 0) + (f12 ? sz[1] : 0) + (f23 ? 0 : sz[3]),
                                 (f03 ? sz[0] : 0) + (f13 ? sz[1] : 0) + (f23 ? sz[2] : 0)};
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int p = base[u] + __popc(bs[u] & lt) + __popc(bs[u + 4] & lt);
                pos[u] = banked_mem_pos(p, K);
                pos[u + 4] = banked_mem_pos(p + (selb[u] ? 1 : 0), K);
            }
        }

        // ---- stage the K entries in shared memory, then coalesced stores --------------------------------
        uint32_t *ent_w = s_ent[warp];
#pragma unroll
        for (int s = 0; s < 8; ++s)
            if (selb[s]) {                    // two 4-byte stores off one address (an 8-byte store needs a register pair)
                uint32_t *e = ent_w + 2 * pos[s];
                e[0] = __float_as_uint(v[s]);
                e[1] = (uint32_t)(8 * lane + s);
            }
        __syncwarp();
#pragma unroll
        for (int i0 = 0; i0 < K; i0 += 32) {
            const int i = i0 + lane;
            if (K >= 32 || i < K) {
                const uint2 ent = *reinterpret_cast<const uint2 *>(ent_w + 2 * i);
                const int c = (int)ent.y;
                const int64_t o = r * K + i;
                if (PLAIN) {
                    out_val[o] = __uint_as_float(ent.x);
                    out_sel[o] = (uint8_t)c;
                } else {
                    if (peers.n == 0) {
                        out_val[o] = __uint_as_float(ent.x);
                        if (out_sel) out_sel[o] = (uint8_t)c;
                    } else if (peers.mc_val != nullptr) {
                        multimem_st_b32(peers.mc_val + o, ent.x);   // the switch replicates the store to every rank
                    } else {
                        for (int p = 0; p < peers.n; ++p) {
                            peers.val[p][o] = __uint_as_float(ent.x);
                            peers.sel[p][o] = (uint8_t)c;
                        }
                    }
                    if (out_i32) out_i32[o] = c;
                    if (out_i64) out_i64[o] = c;
                }
            }
        }
        if (!PLAIN && peers.n != 0 && peers.mc_val != nullptr && lane < K / 4) {   // selectors: 4 per 32-bit multicast store
            uint32_t w = 0u;
#pragma unroll
            for (int b = 0; b < 4; ++b) w |= (ent_w[2 * (4 * lane + b) + 1] & 0xffu) << (8 * b);
            multimem_st_b32(peers.mc_sel + (r * K) / 4 + lane, w);
        }
        __syncwarp();
    }
}

// out[i] = sum over the ranks of the multicast group of their buffers at the same offset, reduced inside the
// NVSwitch (multimem.ld_reduce, SASS LDGMC.ADD): the backward's reduce_scatter of the partial sampled gradient.
__global__ void __launch_bounds__(kTopkThreads)
nvls_reduce_kernel(const float *__restrict__ mc_src, float *__restrict__ dst, int64_t n_vec4)
{
    // 4 independent switch reductions in flight per thread: one at a time left the links at 170 GB/s
    // (tools/exchange_lab.py: 44.6 us for 7.4 MB at 4 GPUs)
    constexpr int U = 4;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < n_vec4; i0 += U * stride) {
        float4 v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t i = i0 + u * stride;
            if (i < n_vec4)
                asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0,%1,%2,%3}, [%4];"
                             : "=f"(v[u].x), "=f"(v[u].y), "=f"(v[u].z), "=f"(v[u].w) : "l"(mc_src + 4 * i) : "memory");
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t i = i0 + u * stride;
            if (i < n_vec4) reinterpret_cast<float4 *>(dst)[i] = v[u];
        }
    }
}

// dense[r, sel[r,l]] = vals[r,l], everything else 0 (one warp per row, row written once).
__global__ void __launch_bounds__(kTopkThreads)
cbsr_scatter_kernel(const float *__restrict__ vals, const uint8_t *__restrict__ sel, int64_t n_rows, int dim, int k,
                    float *__restrict__ dense)
{
    __shared__ __align__(16) float s_row[kTopkWarps][kAccDim];
    const int lane = lane_id();
    const int warp = threadIdx.x >> 5;
    const int64_t warps_total = (int64_t)gridDim.x * kTopkWarps;
    for (int64_t r = (int64_t)blockIdx.x * kTopkWarps + warp; r < n_rows; r += warps_total) {
        for (int j = lane; j < kAccDim; j += 32) s_row[warp][j] = 0.f;
        __syncwarp();
        for (int l = lane; l < k; l += 32) s_row[warp][sel[r * k + l]] = vals[r * k + l];
        __syncwarp();
        float *drow = dense + r * dim;
        if (dim == kAccDim) {
            const float4 *s4 = reinterpret_cast<const float4 *>(s_row[warp]);
            st_stream_f32x4(drow + 4 * lane, s4[lane]);
            st_stream_f32x4(drow + 128 + 4 * lane, s4[32 + lane]);
        } else {
            for (int j = lane; j < dim; j += 32) drow[j] = s_row[warp][j];
        }
        __syncwarp();
    }
}

// out[r,j] = in[r,j] if j selected else 0 (+ add_vals at the selected positions).
__global__ void __launch_bounds__(kTopkThreads)
mask_apply_kernel(const float *__restrict__ in, const uint8_t *__restrict__ sel, const float *__restrict__ add_vals,
                  int64_t n_rows, int dim, int k, float *__restrict__ out)
{
    __shared__ __align__(16) float s_add[kTopkWarps][kAccDim];
    __shared__ uint32_t s_mask[kTopkWarps][8];
    const int lane = lane_id();
    const int warp = threadIdx.x >> 5;
    const int64_t warps_total = (int64_t)gridDim.x * kTopkWarps;
    for (int64_t r = (int64_t)blockIdx.x * kTopkWarps + warp; r < n_rows; r += warps_total) {
        if (lane < 8) s_mask[warp][lane] = 0u;
        if (add_vals)
            for (int j = lane; j < kAccDim; j += 32) s_add[warp][j] = 0.f;
        __syncwarp();
        for (int l = lane; l < k; l += 32) {
            const int c = sel[r * k + l];
            atomicOr(&s_mask[warp][c >> 5], 1u << (c & 31));
            if (add_vals) s_add[warp][c] = add_vals[r * k + l];
        }
        __syncwarp();
        const float *irow = in + r * dim;
        float *orow = out + r * dim;
        for (int j = lane; j < dim; j += 32) {
            const bool on = (s_mask[warp][j >> 5] >> (j & 31)) & 1u;
            float o = on ? irow[j] : 0.f;
            if (add_vals && on) o += s_add[warp][j];
            orow[j] = o;
        }
        __syncwarp();
    }
}

// out = A_csr x dense (validation helper; one warp per row, lanes stride over the columns).
__global__ void __launch_bounds__(kTopkThreads)
dense_spmm_kernel(const int *__restrict__ indptr, const int *__restrict__ idx, const float *__restrict__ val,
                  const float *__restrict__ dense, int64_t n_rows, int dim, float *__restrict__ out)
{
    const int lane = lane_id();
    const int warp = threadIdx.x >> 5;
    const int64_t warps_total = (int64_t)gridDim.x * kTopkWarps;
    for (int64_t r = (int64_t)blockIdx.x * kTopkWarps + warp; r < n_rows; r += warps_total) {
        const int b = indptr[r], e = indptr[r + 1];
        for (int j0 = 0; j0 < dim; j0 += 256) {
            float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            for (int p = b; p < e; ++p) {
                const float w = val[p];
                const float *drow = dense + (size_t)idx[p] * dim + j0;
#pragma unroll
                for (int s = 0; s < 8; ++s) {
                    const int j = lane + 32 * s;
                    if (j0 + j < dim) acc[s] = fmaf(w, drow[j], acc[s]);
                }
            }
#pragma unroll
            for (int s = 0; s < 8; ++s) {
                const int j = j0 + lane + 32 * s;
                if (j < dim) out[r * dim + j] = acc[s];
            }
        }
    }
}

// One resident wave: the grid is what the device holds at once (cached per device), rows beyond it are
// taken by the grid-stride loop, so that no SM idles while a partial second wave drains.
template <int K, bool PLAIN>
static cudaError_t launch_topk_banked(const float *x, int64_t n_rows, float *cbsr_val, uint8_t *cbsr_sel,
                                      int32_t *idx_i32, int64_t *idx_i64, float *masked, cudaStream_t st,
                                      const PeerOut &peers)
{
    static int resident[kMaxCachedDevices];    // CTAs per device, 0 = not queried yet
    int dev = 0;
    cudaError_t err = cudaGetDevice(&dev);
    if (err != cudaSuccess) return err;
    int local = 0;
    int &cap = (dev >= 0 && dev < kMaxCachedDevices) ? resident[dev] : local;
    if (cap == 0) {
        int per_sm = 0;
        err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, topk_banked_kernel<K, PLAIN>, kTopkThreads, 0);
        if (err != cudaSuccess) return err;
        cap = device_sm_count() * (per_sm < 1 ? 1 : per_sm);
    }
    const int64_t need = (n_rows + kTopkWarps - 1) / kTopkWarps;
    const int grid = (int)(need < cap ? need : cap);
    topk_banked_kernel<K, PLAIN><<<grid, kTopkThreads, 0, st>>>(x, n_rows, cbsr_val, cbsr_sel, idx_i32, idx_i64, masked, 1, peers);
    return cudaGetLastError();
}

static cudaError_t dispatch_topk_banked(int k, const float *x, int64_t n_rows, float *cbsr_val, uint8_t *cbsr_sel,
                                        int32_t *idx_i32, int64_t *idx_i64, float *masked, cudaStream_t st,
                                        const PeerOut &peers)
{
    const bool plain = idx_i32 == nullptr && idx_i64 == nullptr && masked == nullptr && peers.n == 0 && cbsr_sel != nullptr;
#define MAXK_TOPK_CASE(KK)                                                                                              \
    case KK:                                                                                                            \
        return plain ? launch_topk_banked<KK, true>(x, n_rows, cbsr_val, cbsr_sel, idx_i32, idx_i64, masked, st, peers) \
                     : launch_topk_banked<KK, false>(x, n_rows, cbsr_val, cbsr_sel, idx_i32, idx_i64, masked, st, peers);
    switch (k) {
        MAXK_TOPK_CASE(8)
        MAXK_TOPK_CASE(16)
        MAXK_TOPK_CASE(32)
        MAXK_TOPK_CASE(64)
        MAXK_TOPK_CASE(96)
        default: break;
    }
    return plain ? launch_topk_banked<128, true>(x, n_rows, cbsr_val, cbsr_sel, idx_i32, idx_i64, masked, st, peers)
                 : launch_topk_banked<128, false>(x, n_rows, cbsr_val, cbsr_sel, idx_i32, idx_i64, masked, st, peers);
#undef MAXK_TOPK_CASE
}

static int grid_for_rows(int64_t n_rows)
{
    int dev = 0, sms = kNumSMsB200;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int64_t need = (n_rows + kTopkWarps - 1) / kTopkWarps;
    const int64_t cap = (int64_t)sms * 8;  // 8 CTAs of 256 threads per SM, grid-stride beyond that
    return (int)(need < cap ? (need > 0 ? need : 1) : cap);
}

}  // namespace maxk

using namespace maxk;

extern "C" int maxk_banked_modulus(int k) { return banked_modulus(k); }

extern "C" int maxk_topk_cbsr(const float *x, int64_t n_rows, int dim, int k, int order, float *cbsr_val,
                              uint8_t *cbsr_sel, int32_t *idx_i32, int64_t *idx_i64, float *masked,
                              maxk_stream_t stream)
{
    if (dim < 1 || dim > kAccDim) return MAXK_ERR_BAD_DIM;
    if (k < 1 || k > dim) return MAXK_ERR_BAD_K;
    if (n_rows < 0) return MAXK_ERR_SIZE;
    if (n_rows == 0) return MAXK_OK;
    if (!x || !cbsr_val) return MAXK_ERR_NULL;
    if (dim == kAccDim && (((uintptr_t)x | (uintptr_t)masked) & 15)) return MAXK_ERR_ALIGN;
    if (order != MAXK_ORDER_VALUE_DESC && order != MAXK_ORDER_COLUMN_ASC && order != MAXK_ORDER_BANKED)
        return MAXK_ERR_SIZE;
    const int grid = grid_for_rows(n_rows);
    const int bm = banked_modulus(k);
    cudaStream_t st = (cudaStream_t)stream;
    if (dim == kAccDim && order == MAXK_ORDER_BANKED && bm >= 4 && !(((uintptr_t)x | (uintptr_t)masked) & 31)) {
        // the layer's hot configurations (32-byte loads need 32-byte aligned rows)
        PeerOut none;
        none.n = 0;
        none.mc_val = nullptr;
        none.mc_sel = nullptr;
        return status_from_cuda(dispatch_topk_banked(k, x, n_rows, cbsr_val, cbsr_sel, idx_i32, idx_i64, masked, st, none));
    }
#define MAXK_TOPK_LAUNCH(D256, ORD) \
    topk_cbsr_kernel<D256, ORD><<<grid, kTopkThreads, 0, st>>>(x, n_rows, dim, k, bm, cbsr_val, cbsr_sel, idx_i32, idx_i64, masked, 1)
    if (dim == kAccDim) {
        if (order == MAXK_ORDER_VALUE_DESC) MAXK_TOPK_LAUNCH(true, MAXK_ORDER_VALUE_DESC);
        else if (order == MAXK_ORDER_COLUMN_ASC || bm < 4) MAXK_TOPK_LAUNCH(true, MAXK_ORDER_COLUMN_ASC);
        else MAXK_TOPK_LAUNCH(true, MAXK_ORDER_BANKED);
    } else {
        if (order == MAXK_ORDER_VALUE_DESC) MAXK_TOPK_LAUNCH(false, MAXK_ORDER_VALUE_DESC);
        else if (order == MAXK_ORDER_COLUMN_ASC || bm < 4) MAXK_TOPK_LAUNCH(false, MAXK_ORDER_COLUMN_ASC);
        else MAXK_TOPK_LAUNCH(false, MAXK_ORDER_BANKED);
    }
#undef MAXK_TOPK_LAUNCH
    return status_from_cuda(cudaGetLastError());
}

extern "C" int maxk_nvls_reduce(const float *mc_src, float *dst, int64_t n_floats, maxk_stream_t stream)
{
    if (n_floats < 0 || (n_floats & 3)) return MAXK_ERR_SIZE;
    if (n_floats == 0) return MAXK_OK;
    if (!mc_src || !dst) return MAXK_ERR_NULL;
    if (((uintptr_t)mc_src | (uintptr_t)dst) & 15) return MAXK_ERR_ALIGN;
    const int64_t n4 = n_floats / 4;
    const int64_t need = (n4 + 4 * kTopkThreads - 1) / (4 * kTopkThreads);
    const int64_t cap = (int64_t)device_sm_count() * 8;
    nvls_reduce_kernel<<<(int)(need < 1 ? 1 : (need < cap ? need : cap)), kTopkThreads, 0, (cudaStream_t)stream>>>(mc_src, dst, n4);
    return status_from_cuda(cudaGetLastError());
}

extern "C" int maxk_topk_cbsr_peers(const float *x, int64_t n_rows, int k, int n_peers, float *const *peer_val,
                                    uint8_t *const *peer_sel, float *mc_val, uint8_t *mc_sel, int64_t row_offset,
                                    float *masked, maxk_stream_t stream)
{
    if (banked_modulus(k) < 4) return MAXK_ERR_BAD_K;
    if (n_rows < 0 || row_offset < 0) return MAXK_ERR_SIZE;
    if (n_peers < 1 || n_peers > MAXK_MAX_PEERS) return MAXK_ERR_SIZE;
    if (n_rows == 0) return MAXK_OK;
    if (!x || !peer_val || !peer_sel) return MAXK_ERR_NULL;
    if (((uintptr_t)x | (uintptr_t)masked) & 31) return MAXK_ERR_ALIGN;
    PeerOut peers;
    peers.n = n_peers;
    for (int p = 0; p < n_peers; ++p) {
        if (!peer_val[p] || !peer_sel[p]) return MAXK_ERR_NULL;
        peers.val[p] = peer_val[p] + row_offset * k;
        peers.sel[p] = peer_sel[p] + row_offset * k;
    }
    const bool mc = mc_val != nullptr && mc_sel != nullptr;
    if (mc && (((uintptr_t)mc_val | (uintptr_t)mc_sel) & 3)) return MAXK_ERR_ALIGN;
    peers.mc_val = mc ? mc_val + row_offset * k : nullptr;
    peers.mc_sel = mc ? reinterpret_cast<uint32_t *>(mc_sel + row_offset * k) : nullptr;
    return status_from_cuda(dispatch_topk_banked(k, x, n_rows, nullptr, nullptr, nullptr, nullptr, masked,
                                                 (cudaStream_t)stream, peers));
}

extern "C" int maxk_cbsr_scatter(const float *vals, const uint8_t *sel, int64_t n_rows, int dim, int k, float *dense,
                                 maxk_stream_t stream)
{
    if (dim < 1 || dim > kAccDim) return MAXK_ERR_BAD_DIM;
    if (k < 1 || k > dim) return MAXK_ERR_BAD_K;
    if (n_rows < 0) return MAXK_ERR_SIZE;
    if (n_rows == 0) return MAXK_OK;
    if (!vals || !sel || !dense) return MAXK_ERR_NULL;
    if (dim == kAccDim && ((uintptr_t)dense & 15)) return MAXK_ERR_ALIGN;
    cbsr_scatter_kernel<<<grid_for_rows(n_rows), kTopkThreads, 0, (cudaStream_t)stream>>>(vals, sel, n_rows, dim, k,
                                                                                          dense);
    return status_from_cuda(cudaGetLastError());
}

extern "C" int maxk_mask_apply(const float *in, const uint8_t *sel, const float *add_vals, int64_t n_rows, int dim,
                               int k, float *out, maxk_stream_t stream)
{
    if (dim < 1 || dim > kAccDim) return MAXK_ERR_BAD_DIM;
    if (k < 1 || k > dim) return MAXK_ERR_BAD_K;
    if (n_rows < 0) return MAXK_ERR_SIZE;
    if (n_rows == 0) return MAXK_OK;
    if (!in || !sel || !out) return MAXK_ERR_NULL;
    mask_apply_kernel<<<grid_for_rows(n_rows), kTopkThreads, 0, (cudaStream_t)stream>>>(in, sel, add_vals, n_rows,
                                                                                        dim, k, out);
    return status_from_cuda(cudaGetLastError());
}

extern "C" int maxk_dense_spmm(const int32_t *indptr, const int32_t *indices, const float *values,
                               const float *dense, int64_t n_rows, int dim, float *out, maxk_stream_t stream)
{
    if (dim < 1) return MAXK_ERR_BAD_DIM;
    if (n_rows < 0) return MAXK_ERR_SIZE;
    if (n_rows == 0) return MAXK_OK;
    if (!indptr || !out) return MAXK_ERR_NULL;
    dense_spmm_kernel<<<grid_for_rows(n_rows), kTopkThreads, 0, (cudaStream_t)stream>>>(indptr, indices, values,
                                                                                        dense, n_rows, dim, out);
    return status_from_cuda(cudaGetLastError());
}
