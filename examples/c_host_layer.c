/*
 * c_host_layer.c -- one MaxK layer pass (top-k -> forward SpGEMM -> backward SSpMM) from a plain C host
 * through include/maxk_b200.h: no Python, no torch, only the CUDA runtime.  This is what a cgo / JNI / Rust
 * binding does, and what replaces main() of the reference's kernels/main.cu:60-190
 * (random graph, random features, launch, read back).
 *
 *   gcc -std=c99 -O2 -Iinclude -I/usr/local/cuda/include examples/c_host_layer.c \
 *       -Lspgemm-prunning_b200/lib -lmaxk_b200 -L/usr/local/cuda/lib64 -lcudart \
 *       -Wl,-rpath,$PWD/spgemm-prunning_b200/lib -o c_host_layer && ./c_host_layer
 *
 * Prints a checksum of the outputs and a CPU re-computation of one output row.
 */
#include <cuda_runtime_api.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "maxk_b200.h"

#define CHECK_CUDA(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)
#define CHECK_MAXK(x) do { int s_ = (x); if (s_ != MAXK_OK) { fprintf(stderr, "%s: %s\n", #x, maxk_status_string(s_)); return 1; } } while (0)

static unsigned int lcg(unsigned int *s) { *s = *s * 1664525u + 1013904223u; return *s >> 8; }

int main(void)
{
    const int n = 4096, deg = 24, dim = 256, k = 32;
    const long long e = (long long)n * deg;
    unsigned int seed = 123;
    int *indptr = (int *)malloc(sizeof(int) * (n + 1));
    int *indices = (int *)malloc(sizeof(int) * e);
    float *values = (float *)malloc(sizeof(float) * e);
    float *x = (float *)malloc(sizeof(float) * n * dim);
    float *grad = (float *)malloc(sizeof(float) * n * dim);
    float *row_div = (float *)malloc(sizeof(float) * n);
    for (int r = 0; r <= n; ++r) indptr[r] = r * deg;
    for (long long i = 0; i < e; ++i) { indices[i] = (int)(lcg(&seed) % n); values[i] = (float)(lcg(&seed) % 1000) / 1000.0f; }
    for (int i = 0; i < n * dim; ++i) { x[i] = (float)(lcg(&seed) % 100003) / 100003.0f; grad[i] = (float)(lcg(&seed) % 1000) / 1000.0f; }
    for (int r = 0; r < n; ++r) row_div[r] = (float)deg;

    int *d_indptr, *d_indices;
    float *d_values, *d_x, *d_grad, *d_div, *d_val, *d_out, *d_gs;
    unsigned char *d_sel;
    void *d_ws;
    const size_t ws_bytes = maxk_spgemm_workspace_bytes(n);
    cudaStream_t stream;
    CHECK_CUDA(cudaStreamCreate(&stream));
    CHECK_CUDA(cudaMalloc((void **)&d_indptr, sizeof(int) * (n + 1)));
    CHECK_CUDA(cudaMalloc((void **)&d_indices, sizeof(int) * e));
    CHECK_CUDA(cudaMalloc((void **)&d_values, sizeof(float) * e));
    CHECK_CUDA(cudaMalloc((void **)&d_x, sizeof(float) * n * dim));
    CHECK_CUDA(cudaMalloc((void **)&d_grad, sizeof(float) * n * dim));
    CHECK_CUDA(cudaMalloc((void **)&d_div, sizeof(float) * n));
    CHECK_CUDA(cudaMalloc((void **)&d_val, sizeof(float) * n * k));
    CHECK_CUDA(cudaMalloc((void **)&d_sel, (size_t)n * k));
    CHECK_CUDA(cudaMalloc((void **)&d_out, sizeof(float) * n * dim));
    CHECK_CUDA(cudaMalloc((void **)&d_gs, sizeof(float) * n * k));
    CHECK_CUDA(cudaMalloc(&d_ws, ws_bytes));
    CHECK_CUDA(cudaMemcpyAsync(d_indptr, indptr, sizeof(int) * (n + 1), cudaMemcpyHostToDevice, stream));
    CHECK_CUDA(cudaMemcpyAsync(d_indices, indices, sizeof(int) * e, cudaMemcpyHostToDevice, stream));
    CHECK_CUDA(cudaMemcpyAsync(d_values, values, sizeof(float) * e, cudaMemcpyHostToDevice, stream));
    CHECK_CUDA(cudaMemcpyAsync(d_x, x, sizeof(float) * n * dim, cudaMemcpyHostToDevice, stream));
    CHECK_CUDA(cudaMemcpyAsync(d_grad, grad, sizeof(float) * n * dim, cudaMemcpyHostToDevice, stream));
    CHECK_CUDA(cudaMemcpyAsync(d_div, row_div, sizeof(float) * n, cudaMemcpyHostToDevice, stream));

    /* the layer: three calls on one stream, nothing synchronises in between */
    CHECK_MAXK(maxk_topk_cbsr(d_x, n, dim, k, MAXK_ORDER_BANKED, d_val, d_sel, NULL, NULL, NULL, (maxk_stream_t)stream));
    CHECK_MAXK(maxk_spgemm_forward(d_indptr, d_indptr + 1, d_indices, d_values, d_val, d_sel, d_out, n, e, dim, k, d_div,
                                   d_ws, ws_bytes, (maxk_stream_t)stream));
    CHECK_MAXK(maxk_sspmm_backward(d_indptr, d_indptr + 1, d_indices, d_values, d_grad, d_sel, d_gs, n, n, e, dim, k, d_div,
                                   d_ws, ws_bytes, (maxk_stream_t)stream));

    float *out = (float *)malloc(sizeof(float) * n * dim);
    float *gs = (float *)malloc(sizeof(float) * n * k);
    float *val = (float *)malloc(sizeof(float) * n * k);
    unsigned char *sel = (unsigned char *)malloc((size_t)n * k);
    CHECK_CUDA(cudaMemcpyAsync(out, d_out, sizeof(float) * n * dim, cudaMemcpyDeviceToHost, stream));
    CHECK_CUDA(cudaMemcpyAsync(gs, d_gs, sizeof(float) * n * k, cudaMemcpyDeviceToHost, stream));
    CHECK_CUDA(cudaMemcpyAsync(val, d_val, sizeof(float) * n * k, cudaMemcpyDeviceToHost, stream));
    CHECK_CUDA(cudaMemcpyAsync(sel, d_sel, (size_t)n * k, cudaMemcpyDeviceToHost, stream));
    CHECK_CUDA(cudaStreamSynchronize(stream));

    /* row 7 of the forward, recomputed on the host from the CBSR the device produced */
    double want[256];
    memset(want, 0, sizeof(want));
    for (int p = indptr[7]; p < indptr[8]; ++p)
        for (int l = 0; l < k; ++l) want[sel[(size_t)indices[p] * k + l]] += (double)values[p] * val[(size_t)indices[p] * k + l];
    double err = 0.0, sum_out = 0.0, sum_gs = 0.0;
    for (int j = 0; j < dim; ++j) err = fmax(err, fabs(want[j] / deg - out[7 * dim + j]));
    for (int i = 0; i < n * dim; ++i) sum_out += out[i];
    for (int i = 0; i < n * k; ++i) sum_gs += gs[i];
    printf("sum(out) = %.6e  sum(gs) = %.6e  max |row 7 - host| = %.3e\n", sum_out, sum_gs, err);
    return err < 1e-4 ? 0 : 2;
}
