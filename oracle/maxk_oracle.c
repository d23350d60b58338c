/*
 * maxk_oracle.c -- CPU restatement of the MaxK aggregation hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (spgemm-prunning_b200/)
 * may import, link or call this file.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs use it, and only as the checker.
 *
 * Parity pinning (see DESIGN.md "Oracle"): the reference ships no golden vectors
 * for this path.  This restatement is pinned against
 *   (1) tests/golden/ref_py_*.npz  -- outputs of the reference's own Python code
 *       (maxk_spgemm_function.py CPU branch, utils/models.py MaxK, kernels/generate_meta.py)
 *       imported from /root/reference by tests/golden/make_golden_py.py, and
 *   (2) tests/golden/ref_cuda_*.npz -- outputs of the reference's own CUDA kernels
 *       (kernels/spmm_maxk.cu, kernels/spmm_maxk_backward.cu) compiled unmodified for
 *       sm_100a into oracle/_ref/ and run on a B200 by tests/golden/make_golden_cuda.py.
 *
 * All accumulation is in double and cast to float once, so the oracle is the
 * "exact" answer both the reference kernels and ours are compared to.
 *
 * Citations are to files under /root/reference.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>

#ifdef _OPENMP
#include <omp.h>
#endif

/* ---------------------------------------------------------------------------
 * Order-preserving key for "value descending, NaN largest, -0 == +0".
 * Semantic source: torch.topk(x, k, dim=1) as called at
 *   maxk_spgemm_function.py:53, model_integrated_v3.py:32, spgemmfunction_v4:50-51.
 * torch orders NaN above +inf; ties are broken here by LOWEST column (north star),
 * which torch.topk does not promise -- the golden test therefore compares with
 * torch only on tie-free rows and checks tie rows against this definition.
 * ------------------------------------------------------------------------- */
static inline uint32_t order_key(float f)
{
    uint32_t b;
    memcpy(&b, &f, 4);
    if ((b & 0x7fffffffu) > 0x7f800000u) return 0xffffffffu; /* NaN: largest   */
    if (b == 0x80000000u) b = 0;                              /* -0 -> +0       */
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

typedef struct { uint32_t key; int32_t col; } kc_t;

/* descending key, ascending column */
static int kc_cmp(const void *a, const void *b)
{
    const kc_t *x = (const kc_t *)a, *y = (const kc_t *)b;
    if (x->key != y->key) return (x->key < y->key) ? 1 : -1;
    return (x->col > y->col) - (x->col < y->col);
}

/*
 * MaxK top-k -> CBSR.
 *   x        [N, D] fp32 row-major
 *   out_val  [N, k] fp32   k largest of the row
 *   out_sel  [N, k] int32  their columns
 * order = 0: entries sorted (value desc, column asc)   -- torch.topk(sorted=True) order
 * order = 1: entries sorted by column ascending
 * order = 2: MAXK_ORDER_BANKED, the order our fused kernels emit (a pure permutation of the same k entries;
 *            consumers are order independent).  With bank_mod = m > 1: residue classes column mod m, the
 *            class with the most entries first (ties: lower class), columns ascending inside a class; the
 *            entry of sorted rank p is stored at position p for k < 32 and, for k >= 32, at
 *            32 * (i / 8) + 8 * t + i % 8 with t = p / (k/4), i = p % (k/4) (lane t of a 4-lane slot owns k/4
 *            consecutive ranks, kept as 8-entry runs per 32-entry chunk).  bank_mod <= 1: column order.
 */
void oracle_topk(const float *x, int64_t N, int D, int k, float *out_val, int32_t *out_sel, int order, int bank_mod)
{
#pragma omp parallel
    {
        kc_t *buf = (kc_t *)malloc(sizeof(kc_t) * (size_t)D);
        kc_t *tmp = (kc_t *)malloc(sizeof(kc_t) * (size_t)D);
#pragma omp for schedule(static)
        for (int64_t r = 0; r < N; ++r) {
            const float *row = x + r * D;
            for (int j = 0; j < D; ++j) { buf[j].key = order_key(row[j]); buf[j].col = j; }
            qsort(buf, (size_t)D, sizeof(kc_t), kc_cmp);
            if (order == 1 || order == 2) {
                const int m = (order == 2 && bank_mod > 1) ? bank_mod : 1;
                int size[64] = {0}, rank[64];
                for (int i = 0; i < k; ++i) size[buf[i].col % m]++;
                for (int u = 0; u < m; ++u) {              /* classes by (size desc, class asc) */
                    rank[u] = 0;
                    for (int v = 0; v < m; ++v)
                        if (v != u && (size[v] > size[u] || (size[v] == size[u] && v < u))) rank[u]++;
                }
                /* sort the chosen k by (class rank, col): simple insertion sort */
                for (int i = 1; i < k; ++i) {
                    kc_t t = buf[i]; int j = i - 1;
                    while (j >= 0 && (rank[buf[j].col % m] > rank[t.col % m] ||
                                      (buf[j].col % m == t.col % m && buf[j].col > t.col))) { buf[j + 1] = buf[j]; --j; }
                    buf[j + 1] = t;
                }
                if (m > 1 && k >= 32) {                    /* rank -> position in the row */
                    const int epl = k / 4;
                    for (int p = 0; p < k; ++p) {
                        const int t = p / epl, i = p % epl;
                        tmp[32 * (i / 8) + 8 * t + i % 8] = buf[p];
                    }
                    for (int p = 0; p < k; ++p) buf[p] = tmp[p];
                }
            }
            for (int i = 0; i < k; ++i) {
                out_val[r * k + i] = row[buf[i].col];
                out_sel[r * k + i] = buf[i].col;
            }
        }
        free(buf);
        free(tmp);
    }
}

/*
 * warp4 partition metadata, restating kernels/generate_meta.py:30-48:
 * every non-empty CSR row r yields ceil(deg/max_nz) quads (row, loc, len<=max_nz, 0).
 * Returns the number of quads W; when out == NULL only counts.
 */
int64_t oracle_warp4(const int32_t *indptr, int64_t N, int max_nz, int32_t *out)
{
    int64_t w = 0;
    for (int64_t r = 0; r < N; ++r) {
        int32_t beg = indptr[r], end = indptr[r + 1];
        for (int32_t loc = beg; loc < end; loc += max_nz) {      /* generate_meta.py:35-45 */
            if (out) {
                int32_t len = end - loc < max_nz ? end - loc : max_nz;
                out[4 * w + 0] = (int32_t)r;
                out[4 * w + 1] = loc;
                out[4 * w + 2] = len;
                out[4 * w + 3] = 0;                               /* generate_meta.py:46 */
            }
            ++w;
        }
    }
    return w;
}

/*
 * Forward row-wise-product SpGEMM, restating kernels/spmm_maxk.cu:62-105:
 *   out[r, sel[c,l]] += val[e] * data[c,l]   for every edge e=(r,c), l<k
 * CSR given by indptr/idx/val over n_rows rows; CBSR (data, sel) over the source nodes.
 * deg (nullable): fp32 divisor applied AFTER the sum exactly as the Python layer does
 *   (maxk_spgemm_function.py:86, spgemmfunction_v4:72): out32 = float(sum); out32 /= deg[r].
 */
/* selector width: the reference's format is uint8 (kernels/spmm_maxk.cu:17); the wide path (dim > 256)
 * stores uint16 -- same loops, wider index */
static inline int sel_get(const void *sel, int wide, int64_t i)
{
    return wide ? (int)((const uint16_t *)sel)[i] : (int)((const uint8_t *)sel)[i];
}

static void spgemm_fwd_impl(const int32_t *indptr, const int32_t *idx, const float *val,
                            const float *data, const void *sel, int wide,
                            int64_t n_rows, int k, int D, const float *deg, float *out)
{
#pragma omp parallel
    {
        double *acc = (double *)malloc(sizeof(double) * (size_t)D);
#pragma omp for schedule(dynamic, 64)
        for (int64_t r = 0; r < n_rows; ++r) {
            for (int j = 0; j < D; ++j) acc[j] = 0.0;
            for (int32_t e = indptr[r]; e < indptr[r + 1]; ++e) {
                int64_t c = idx[e];
                double w = (double)val[e];
                const float *dv = data + c * k;
                for (int l = 0; l < k; ++l) acc[sel_get(sel, wide, c * k + l)] += w * (double)dv[l];
            }
            for (int j = 0; j < D; ++j) {
                float o = (float)acc[j];
                if (deg) o = o / deg[r];
                out[r * D + j] = o;
            }
        }
        free(acc);
    }
}

void oracle_spgemm_fwd(const int32_t *indptr, const int32_t *idx, const float *val,
                       const float *data, const uint8_t *sel,
                       int64_t n_rows, int k, int D, const float *deg, float *out)
{
    spgemm_fwd_impl(indptr, idx, val, data, sel, 0, n_rows, k, D, deg, out);
}

void oracle_spgemm_fwd16(const int32_t *indptr, const int32_t *idx, const float *val,
                         const float *data, const uint16_t *sel,
                         int64_t n_rows, int k, int D, const float *deg, float *out)
{
    spgemm_fwd_impl(indptr, idx, val, data, sel, 1, n_rows, k, D, deg, out);
}

/*
 * Same contraction driven by warp4 quads, the way the reference kernel consumes them
 * (kernels/spmm_maxk.cu:37-47 quad decode, :85-96 edge loop, :101-105 per-segment flush
 * added into the zero-filled output of cuda_kernel_bindings.cpp:71).  Rows that appear
 * in no quad stay 0.  out64 is an [n_rows*D] double scratch supplied by the caller.
 */
void oracle_spgemm_fwd_warp4(const int32_t *warp4, int64_t W, const int32_t *idx, const float *val,
                             const float *data, const uint8_t *sel,
                             int64_t n_rows, int k, int D, double *out64, float *out)
{
    memset(out64, 0, sizeof(double) * (size_t)(n_rows * D));
    for (int64_t q = 0; q < W; ++q) {
        int64_t r = warp4[4 * q]; int32_t loc = warp4[4 * q + 1], len = warp4[4 * q + 2];
        for (int32_t e = loc; e < loc + len; ++e) {
            int64_t c = idx[e];
            double w = (double)val[e];
            for (int l = 0; l < k; ++l)
                out64[r * D + sel[c * k + l]] += w * (double)data[c * k + l];
        }
    }
    for (int64_t i = 0; i < n_rows * D; ++i) out[i] = (float)out64[i];
}

/*
 * Backward outer-product SSpMM, restating kernels/spmm_maxk_backward.cu:52-113:
 *   gs[c, l] += val[e] * g[r, sel[c,l]]      for every edge e=(r,c), l<k
 * i.e. gs = sample_sel(A^T g).  n_rows = rows of the CSR / of g; n_cols = rows of sel / gs.
 * deg (nullable): fp32 divisor applied to g BEFORE the product, as the Python layer does
 *   (maxk_spgemm_function.py:155, spgemmfunction_v4:87): g32 = g[r,:] / deg[r].
 * Done destination-major through a counting-sort transpose so it is parallel and
 * order-deterministic.
 */
static void sspmm_bwd_impl(const int32_t *indptr, const int32_t *idx, const float *val,
                           const float *g, const void *sel, int wide,
                           int64_t n_rows, int64_t n_cols, int k, int D, const float *deg, float *gs)
{
    int64_t E = indptr[n_rows];
    int64_t *cptr = (int64_t *)calloc((size_t)n_cols + 1, sizeof(int64_t));
    int32_t *crow = (int32_t *)malloc(sizeof(int32_t) * (size_t)(E > 0 ? E : 1));
    float *cval = (float *)malloc(sizeof(float) * (size_t)(E > 0 ? E : 1));
    for (int64_t e = 0; e < E; ++e) cptr[idx[e] + 1]++;
    for (int64_t c = 0; c < n_cols; ++c) cptr[c + 1] += cptr[c];
    int64_t *fill = (int64_t *)malloc(sizeof(int64_t) * (size_t)(n_cols > 0 ? n_cols : 1));
    memcpy(fill, cptr, sizeof(int64_t) * (size_t)n_cols);
    for (int64_t r = 0; r < n_rows; ++r)
        for (int32_t e = indptr[r]; e < indptr[r + 1]; ++e) {
            int64_t p = fill[idx[e]]++;
            crow[p] = (int32_t)r; cval[p] = val[e];
        }
#pragma omp parallel for schedule(dynamic, 64)
    for (int64_t c = 0; c < n_cols; ++c) {
        for (int l = 0; l < k; ++l) {
            int s = sel_get(sel, wide, c * k + l);
            double acc = 0.0;
            for (int64_t p = cptr[c]; p < cptr[c + 1]; ++p) {
                int64_t r = crow[p];
                float gv = g[r * D + s];
                if (deg) gv = gv / deg[r];
                acc += (double)cval[p] * (double)gv;
            }
            gs[c * k + l] = (float)acc;
        }
    }
    free(cptr); free(crow); free(cval); free(fill);
}

void oracle_sspmm_bwd(const int32_t *indptr, const int32_t *idx, const float *val,
                      const float *g, const uint8_t *sel,
                      int64_t n_rows, int64_t n_cols, int k, int D, const float *deg, float *gs)
{
    sspmm_bwd_impl(indptr, idx, val, g, sel, 0, n_rows, n_cols, k, D, deg, gs);
}

void oracle_sspmm_bwd16(const int32_t *indptr, const int32_t *idx, const float *val,
                        const float *g, const uint16_t *sel,
                        int64_t n_rows, int64_t n_cols, int k, int D, const float *deg, float *gs)
{
    sspmm_bwd_impl(indptr, idx, val, g, sel, 1, n_rows, n_cols, k, D, deg, gs);
}

/*
 * MaxK nonlinearity forward, restating maxk_models_integrated.py:28-37 /
 * utils/models.py:11-20: mask = scatter(topk idx, 1); out = x * mask.
 * sel is the [N,k] column list from oracle_topk.
 */
void oracle_maxk_act_fwd(const float *x, const int32_t *sel, int64_t N, int D, int k, float *out)
{
#pragma omp parallel for schedule(static)
    for (int64_t r = 0; r < N; ++r) {
        for (int j = 0; j < D; ++j) out[r * D + j] = 0.0f; /* x*0: equal to the reference for finite x (-0 == 0) */
        for (int l = 0; l < k; ++l) { int c = sel[r * k + l]; out[r * D + c] = x[r * D + c]; }
    }
}

/* MaxK backward: grad * mask (maxk_models_integrated.py:40-43). */
void oracle_maxk_act_bwd(const float *grad, const int32_t *sel, int64_t N, int D, int k, float *out)
{
#pragma omp parallel for schedule(static)
    for (int64_t r = 0; r < N; ++r) {
        for (int j = 0; j < D; ++j) out[r * D + j] = 0.0f;
        for (int l = 0; l < k; ++l) { int c = sel[r * k + l]; out[r * D + c] = grad[r * D + c]; }
    }
}

int oracle_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
