/*
 * maxk_b200.h -- C ABI of the B200-native MaxK aggregation hot path.
 *
 * This header is the drop-in boundary.  Every entry point takes plain device
 * pointers and sizes (no torch types), launches on the caller's stream, never
 * synchronises, never prints and returns a status (0 = ok).  All pointers are
 * DEVICE pointers unless a parameter says "host".  Outputs are written in full
 * by the callee (no caller-side zero fill is required).
 *
 * Each function cites the reference interface it replaces (paths are relative to
 * julius-sk/spgemm-prunning).  The reference's own C ABI
 * (cuda_kernel_wrappers.cu:36-93, declared cuda_kernel_bindings.cpp:11-35) passes
 * dim3 by value through extern "C" and launches on the legacy stream; this ABI is
 * the same three operations with a real C signature.
 *
 * Layouts
 *   CSR adjacency : row_begin[n_rows], row_end[n_rows] int32 (for a plain CSR pass
 *                   indptr and indptr+1), indices[E] int32 source-node ids, values[E] fp32.
 *   CBSR features : cbsr_val[n_src, k] fp32 + cbsr_sel[n_src, k] uint8 column ids,
 *                   row stride exactly k, columns distinct within a row
 *                   (kernels/spmm_maxk.cu:17 vin_data / vin_selector).
 *   warp4         : int32 quads (row, loc, len<=max_nz, 0) (kernels/generate_meta.py:30-48).
 */
#ifndef MAXK_B200_H
#define MAXK_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Status codes: 0 ok, >0 a cudaError_t from the launch, <0 argument errors below. */
#define MAXK_OK 0
#define MAXK_ERR_BAD_K (-1)          /* k < 1 or k > dim (top-k) / k > 256 (SpGEMM, SSpMM)  */
#define MAXK_ERR_BAD_DIM (-2)        /* dim < 1 or dim > 256 (uint8 selector)  */
#define MAXK_ERR_NULL (-3)           /* a required pointer is NULL                         */
#define MAXK_ERR_WORKSPACE (-4)      /* workspace too small                                */
#define MAXK_ERR_ALIGN (-5)          /* a pointer is not 16-byte aligned                   */
#define MAXK_ERR_SIZE (-6)           /* negative or overflowing size                       */

typedef void *maxk_stream_t; /* a cudaStream_t */

/* Library identification; also lets a binding check it loaded the right .so. */
int maxk_abi_version(void);
const char *maxk_status_string(int status);

/* top-k output order */
#define MAXK_ORDER_VALUE_DESC 0 /* (value desc, column asc): torch.topk(sorted=True) order */
#define MAXK_ORDER_COLUMN_ASC 1 /* column ascending                                        */
#define MAXK_ORDER_BANKED 2     /* the order on which (2) has the fewest shared-memory bank conflicts; the fused
                                   layer uses it.  m = maxk_banked_modulus(k).  m == 1: column ascending.
                                   m == 4 (k in {8,16,32,64,96,128}): residue classes column mod 4, the class
                                   with the most entries first (ties: lower class), columns ascending inside
                                   a class; the entry of sorted rank p is stored at position p for k < 32 and
                                   at 32*(i/8) + 8*t + i%8, t = p/(k/4), i = p%(k/4), for k >= 32 (lane t of a
                                   4-lane slot owns k/4 consecutive ranks as 8-entry runs).  Every consumer
                                   accepts any entry order; only speed depends on it.             */

/*
 * (1) MaxK row-wise top-k -> CBSR.
 * Replaces torch.topk(x, k, dim=1) + .to(uint8) (maxk_spgemm_function.py:53-57,
 * model_integrated_v3.py:32, spgemmfunction_v4:50-51) and the uint8 `topk` kernel behind
 * cuda_topk_maxk / cuda_topk_maxk_float / prepare_cbsr_format_maxk
 * (cuda_kernel_bindings.cpp:164-251, kernels/maxk_kernel.cu:23-96).
 * Exact: the k largest of each row, NaN largest, -0 == +0, ties broken by LOWEST column.
 *   x          [n_rows, dim] fp32
 *   cbsr_val   [n_rows, k]   fp32   (required)
 *   cbsr_sel   [n_rows, k]   uint8  (nullable)
 *   idx_i32    [n_rows, k]   int32  (nullable; the dtype cuda_topk_maxk_float returns)
 *   idx_i64    [n_rows, k]   int64  (nullable; the dtype torch.topk returns)
 *   masked     [n_rows, dim] fp32   (nullable; x with every non-selected entry set to 0 =
 *                                    the MaxK nonlinearity output, maxk_models_integrated.py:28-37)
 * For dim == 256, x and masked must be 16-byte aligned (MAXK_ERR_ALIGN otherwise); when they are 32-byte
 * aligned (any torch allocation) and order == MAXK_ORDER_BANKED with k in {8, 16, 32, 64, 96, 128}, the
 * specialised kernel of the fused layer runs (32-byte loads), else the general one.  Same results.
 */
int maxk_banked_modulus(int k); /* 4 for k in {8, 16, 32, 64, 96, 128}; 1 otherwise */
int maxk_topk_cbsr(const float *x, int64_t n_rows, int dim, int k, int order,
                   float *cbsr_val, uint8_t *cbsr_sel, int32_t *idx_i32, int64_t *idx_i64,
                   float *masked, maxk_stream_t stream);

/*
 * (1b) The same top-k for one rank's row slab of the row-sharded layer (the reference is single-GPU,
 * all_train.py:224): row r of x becomes row (row_offset + r) of EVERY peer's gathered CBSR buffers
 * peer_val[p] [P*m, k] / peer_sel[p] [P*m, k] (host arrays of n_peers <= MAXK_MAX_PEERS device pointers, the
 * rank's own buffer included; peer-mapped memory, e.g. torch symmetric memory or cudaIpc mappings).  Replaces
 * local top-k + two all_gather launches; the caller synchronises the ranks afterwards (one barrier).
 * dim is 256, order MAXK_ORDER_BANKED, k in {8, 16, 32, 64, 96, 128}; x (and masked, nullable) 32-byte aligned.
 * mc_val / mc_sel (nullable): NVLS multicast mappings of the same two buffers; when given, every row is written
 * with ONE multimem.st per 4 bytes and the NVSwitch replicates it to all ranks (egress 5k bytes per node instead
 * of P times that).
 * maxk_nvls_reduce: dst[i] = sum over the ranks of the multicast group of src_rank[i] (multimem.ld_reduce, the
 * sum is formed inside the switch): the backward's reduce_scatter of the partial sampled gradients, each rank
 * reducing its own slab.  mc_src is a multicast address, n_floats a multiple of 4, both pointers 16-byte aligned;
 * the caller synchronises the ranks before (all partials complete) as for any collective.
 */
#define MAXK_MAX_PEERS 8
int maxk_topk_cbsr_peers(const float *x, int64_t n_rows, int k, int n_peers,
                         float *const *peer_val, uint8_t *const *peer_sel,
                         float *mc_val, uint8_t *mc_sel, int64_t row_offset,
                         float *masked, maxk_stream_t stream);
int maxk_nvls_reduce(const float *mc_src, float *dst, int64_t n_floats, maxk_stream_t stream);

/*
 * (2) Forward row-wise-product SpGEMM: out = A_csr x scatter(CBSR)   (optionally / row_div).
 * Replaces spmm_kernel_opt2_sparse_v3_wrapper / spmm_maxk_forward
 * (cuda_kernel_wrappers.cu:38-56, cuda_kernel_bindings.cpp:42-104, kernels/spmm_maxk.cu:17-106).
 *   out[r, cbsr_sel[c,l]] += values[e] * cbsr_val[c,l]  for every edge e=(r,c), l<k
 *   out      [n_rows, dim] fp32, fully written (rows without edges become 0)
 *   row_div  [n_rows] fp32 nullable: out[r,:] /= row_div[r] fused into the epilogue
 *            (the reference does it in Python: maxk_spgemm_function.py:86, spgemmfunction_v4:72)
 *   workspace: maxk_spgemm_workspace_bytes(n_rows) bytes of device scratch, 16-byte aligned.
 * out must be 32-byte aligned when dim == 256.  Deterministic: each output row is reduced in a fixed order
 * (no atomics).  k in {8, 16, 32, 64, 96, 128} with 32-byte aligned cbsr_val / 8-byte aligned cbsr_sel run the
 * vectorised kernels, every other k in [1, 256] a scalar-load variant of the same scheme.
 *
 * The kernel walks the rows in the order of a ROW PLAN (rows sorted by degree bucket and cut into work items,
 * the GPU-built counterpart of the reference's warp4 partitioning, kernels/generate_meta.py:30-48).
 * maxk_spgemm_forward builds the plan into the workspace on every call (three small kernels, no host
 * read-back); a caller that runs many layers / epochs on one graph builds it once with maxk_plan_build and
 * calls maxk_spgemm_forward_planned:
 *   plan       maxk_plan_bytes(n_rows) bytes, 16-byte aligned; a pure function of (row_begin, row_end, device).
 *              Its last 256 bytes are the scheduler tickets of the launches that use it (each launch takes the
 *              next of 32 slots and leaves it zeroed), so the buffer must stay writable and on its device.
 *   workspace  maxk_plan_workspace_bytes(n_rows) bytes of scratch, free again when the build has run
 */
int maxk_spgemm_forward(const int32_t *row_begin, const int32_t *row_end,
                        const int32_t *indices, const float *values,
                        const float *cbsr_val, const uint8_t *cbsr_sel,
                        float *out, int64_t n_rows, int64_t n_edges, int dim, int k,
                        const float *row_div,
                        void *workspace, size_t workspace_bytes, maxk_stream_t stream);
size_t maxk_plan_bytes(int64_t n_rows);
size_t maxk_plan_workspace_bytes(int64_t n_rows);
int maxk_plan_build(const int32_t *row_begin, const int32_t *row_end, int64_t n_rows,
                    void *plan, size_t plan_bytes, void *workspace, size_t workspace_bytes,
                    maxk_stream_t stream);
int maxk_spgemm_forward_planned(const void *plan, const int32_t *indices, const float *values,
                                const float *cbsr_val, const uint8_t *cbsr_sel,
                                float *out, int64_t n_rows, int64_t n_edges, int dim, int k,
                                const float *row_div, maxk_stream_t stream);

/*
 * (3) Backward outer-product SSpMM: gs = sample_sel(A^T (g / row_div)).
 * Replaces spmm_kernel_opt2_sparse_backward_v3_wrapper / spmm_maxk_backward
 * (cuda_kernel_wrappers.cu:58-76, cuda_kernel_bindings.cpp:106-161,
 *  kernels/spmm_maxk_backward.cu:15-115).
 *   gs[c, l] += values[e] * g[r, cbsr_sel[c,l]]          for every edge e=(r,c), l<k
 *   g        [n_rows, dim] fp32   gs [n_dst, k] fp32, fully written (zero-filled inside)
 *   row_div  [n_rows] fp32 nullable: g[r,:] / row_div[r] fused into the row stage
 *            (maxk_spgemm_function.py:155, spgemmfunction_v4:87)
 * Called with the same CSR as (2) it is the exact adjoint of (2).
 */
int maxk_sspmm_backward(const int32_t *row_begin, const int32_t *row_end,
                        const int32_t *indices, const float *values,
                        const float *g, const uint8_t *cbsr_sel,
                        float *gs, int64_t n_rows, int64_t n_dst, int64_t n_edges, int dim, int k,
                        const float *row_div,
                        void *workspace, size_t workspace_bytes, maxk_stream_t stream);

/*
 * Same as maxk_sspmm_backward but gs is NOT zero-filled: gs += the contribution of these rows.
 * Lets a caller feed the source rows in slabs (e.g. while later slabs of g are still in flight
 * from the host) -- zero gs once, then one call per slab with offset row pointers.
 */
int maxk_sspmm_backward_accumulate(const int32_t *row_begin, const int32_t *row_end,
                                   const int32_t *indices, const float *values,
                                   const float *g, const uint8_t *cbsr_sel,
                                   float *gs, int64_t n_rows, int64_t n_dst, int64_t n_edges, int dim, int k,
                                   const float *row_div,
                                   void *workspace, size_t workspace_bytes, maxk_stream_t stream);

/* Scratch needed by (2) and (3) for a CSR with n_rows rows. */
size_t maxk_spgemm_workspace_bytes(int64_t n_rows);

/*
 * (4) warp4 partition metadata built on the GPU.
 * Replaces the host loop of kernels/generate_meta.py:30-48 and the file round trip of
 * load_warp4_metadata (cuda_kernel_bindings.cpp:287-317).
 * Two phases because the quad count W must be known to allocate:
 *   maxk_warp4_scan : seg_offsets[n_rows+1] int32 = exclusive scan of ceil(deg/max_nz);
 *                     W = seg_offsets[n_rows] (read it back on the host).
 *   maxk_warp4_fill : warp4[4*W] from indptr + seg_offsets.
 *   scan_workspace  : maxk_warp4_workspace_bytes(n_rows) bytes.
 */
size_t maxk_warp4_workspace_bytes(int64_t n_rows);
int maxk_warp4_scan(const int32_t *indptr, int64_t n_rows, int max_nz, int32_t *seg_offsets,
                    void *workspace, size_t workspace_bytes, maxk_stream_t stream);
int maxk_warp4_fill(const int32_t *indptr, const int32_t *seg_offsets, int64_t n_rows, int max_nz,
                    int32_t *warp4, maxk_stream_t stream);
/*
 * Inverse: recover per-row edge ranges from warp4 quads so that (2)/(3) can be driven by
 * the reference's metadata alone (spmm_maxk_forward receives warp4 but no indptr,
 * cuda_kernel_bindings.cpp:42-50).  Rows absent from warp4 get begin = end = 0.
 * Quads must be well formed (generate_meta.py): grouped by row, contiguous locs.
 */
int maxk_warp4_to_rows(const int32_t *warp4, int64_t n_quads, int64_t n_rows,
                       int32_t *row_begin, int32_t *row_end, maxk_stream_t stream);

/*
 * (6) Feature widths above 256: CBSR with uint16 column selectors, dim <= 1024.
 * The reference cannot address them: it casts torch.topk's indices to uint8 (maxk_spgemm_function.py:57) and
 * hard-wires 256 output columns (cuda_kernel_bindings.cpp:70), although its own Yelp script trains with hidden 384
 * (scripts_train/yelp_maxk.sh:16).  Same contracts as (1)-(3) with
 *   cbsr_sel  [n, k] uint16, entries in column-ascending order, k <= 256
 *   out / g   [n_rows, dim] fp32 dense rows of dim <= 1024 floats (out 32-byte aligned, dim % 8 == 0 for the vector stores)
 * The forward runs on a row plan (maxk_plan_build); both operators need
 * maxk_wide_workspace_bytes(n_rows, n_src, dim, k) bytes of 256-byte aligned scratch (n_src = rows of the CBSR).
 */
int maxk_topk_cbsr16(const float *x, int64_t n_rows, int dim, int k,
                     float *cbsr_val, uint16_t *cbsr_sel, float *masked, maxk_stream_t stream);
size_t maxk_wide_workspace_bytes(int64_t n_rows, int64_t n_src, int dim, int k);
int maxk_spgemm_forward16(const void *plan, const int32_t *indices, const float *values,
                          const float *cbsr_val, const uint16_t *cbsr_sel,
                          float *out, int64_t n_rows, int64_t n_src, int64_t n_edges, int dim, int k,
                          const float *row_div, void *workspace, size_t workspace_bytes, maxk_stream_t stream);
int maxk_sspmm_backward16(const int32_t *row_begin, const int32_t *row_end,
                          const int32_t *indices, const float *values,
                          const float *g, const uint16_t *cbsr_sel,
                          float *gs, int64_t n_rows, int64_t n_dst, int64_t n_edges, int dim, int k,
                          const float *row_div, void *workspace, size_t workspace_bytes, maxk_stream_t stream);

/*
 * (5) Helpers of the operator surface.
 * maxk_cbsr_scatter: dense[r, sel[r,l]] = vals[r,l], everything else 0.  Replaces
 *   zeros(...).scatter_(1, sel.long(), grad_sparse) (maxk_spgemm_function.py:152,175).
 * maxk_mask_apply  : out[r,j] = in[r,j] if j in sel[r,:] else 0.  Replaces grad * mask of
 *   MaxK.backward (maxk_models_integrated.py:40-43) without the saved fp32 mask.
 *   When add_vals is non-NULL, out[r, sel[r,l]] additionally += add_vals[r,l]
 *   (the aggregation gradient that OPTMaxK.backward drops).
 * maxk_dense_spmm  : out = A_csr x dense, fp32.  Stands in for the reference's
 *   cusparse_spmm validation helper (cuda_kernel_bindings.cpp:253-284) without cuSPARSE.
 */
int maxk_cbsr_scatter(const float *vals, const uint8_t *sel, int64_t n_rows, int dim, int k,
                      float *dense, maxk_stream_t stream);
int maxk_mask_apply(const float *in, const uint8_t *sel, const float *add_vals,
                    int64_t n_rows, int dim, int k, float *out, maxk_stream_t stream);
int maxk_dense_spmm(const int32_t *indptr, const int32_t *indices, const float *values,
                    const float *dense, int64_t n_rows, int dim, float *out, maxk_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* MAXK_B200_H */
