// Microbenchmark (developer tool): does staging the gathered CBSR rows in shared memory with the bulk async
// copy engine (cp.async.bulk, "TMA without a tensor map") beat register gathers (LDG) on B200?
// north_star (2) names "TMA or cp.async staging of neighbour rows"; this measures it.
//
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/_bin/bulk_gather tools/bulk_gather.cu
//
// Workload: 48 Mi random gathers of k=32 CBSR rows (128 B values + 32 B selectors) out of a 232,965-row table
// (37 MB, L2 resident) -- the Reddit-shape forward's gather stream.  Every variant consumes the rows the way the
// SpGEMM does (each row is read once from where it landed and reduced into a checksum).
//   ldg      : 4 lanes x 32 B values + 4 lanes x 8 B selectors per row, 8 rows per warp instruction (registers)
//   bulk     : every lane issues cp.async.bulk for one row (128 B + 32 B) into a per-warp shared-memory ring
//              (STAGES x 32 rows), completion on an mbarrier; rows are then read back with LDS.128 / LDS.64
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(256) gather_ldg(const float *__restrict__ vals, const unsigned char *__restrict__ sel,
                                                  const int *__restrict__ idx, size_t nidx, float *out)
{
    const size_t warp = (blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5, nwarps = ((size_t)gridDim.x * blockDim.x) >> 5;
    const int lane = threadIdx.x & 31, q = lane >> 2, t = lane & 3;
    float acc = 0.f;
    for (size_t i = warp * 32; i + 32 <= nidx; i += nwarps * 32) {
        const int mine = __ldg(idx + i + lane);
        float v[4][8];
        uint2 s[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int c = __shfl_sync(0xffffffffu, mine, 8 * u + q);
            const float *p = vals + (size_t)c * 32 + 8 * t;
            asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                         : "=f"(v[u][0]), "=f"(v[u][1]), "=f"(v[u][2]), "=f"(v[u][3]), "=f"(v[u][4]), "=f"(v[u][5]),
                           "=f"(v[u][6]), "=f"(v[u][7]) : "l"(p));
            s[u] = __ldg(reinterpret_cast<const uint2 *>(sel + (size_t)c * 32) + t);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
#pragma unroll
            for (int e = 0; e < 8; ++e) acc += v[u][e];
            acc += (float)(s[u].x ^ s[u].y);
        }
    }
    if (acc == 12345.f) out[0] = acc;
}

template <int STAGES>
__global__ void __launch_bounds__(256) gather_bulk(const float *__restrict__ vals, const unsigned char *__restrict__ sel,
                                                   const int *__restrict__ idx, size_t nidx, float *out)
{
    extern __shared__ __align__(128) unsigned char smem[];
    const int lane = threadIdx.x & 31, warp_in = threadIdx.x >> 5, q = lane >> 2, t = lane & 3;
    const size_t warp = (blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5, nwarps = ((size_t)gridDim.x * blockDim.x) >> 5;
    constexpr int kStageBytes = 32 * 160;
    unsigned char *ring = smem + (size_t)warp_in * (STAGES * kStageBytes + 64);
    uint64_t *bars = reinterpret_cast<uint64_t *>(ring + STAGES * kStageBytes);
    if (lane == 0)
        for (int s = 0; s < STAGES; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bars + s)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncwarp();
    float acc = 0.f;
    const size_t step = nwarps * 32;
    size_t n_batches = 0;
    for (size_t i = warp * 32; i + 32 <= nidx; i += step) ++n_batches;
    auto issue = [&](size_t b) {
        const int s = (int)(b % STAGES);
        const int c = __ldg(idx + warp * 32 + b * step + lane);
        if (lane == 0)
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bars + s)), "r"(kStageBytes) : "memory");
        __syncwarp();
        unsigned char *dst = ring + s * kStageBytes;
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], 128, [%2];" ::"r"(
                         smem_u32(dst + lane * 128)), "l"(vals + (size_t)c * 32), "r"(smem_u32(bars + s)) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], 32, [%2];" ::"r"(
                         smem_u32(dst + 32 * 128 + lane * 32)), "l"(sel + (size_t)c * 32), "r"(smem_u32(bars + s)) : "memory");
    };
    for (size_t b = 0; b < STAGES - 1 && b < n_batches; ++b) issue(b);
    for (size_t b = 0; b < n_batches; ++b) {
        if (b + STAGES - 1 < n_batches) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            issue(b + STAGES - 1);
        }
        const int s = (int)(b % STAGES);
        const uint32_t parity = (uint32_t)((b / STAGES) & 1);
        uint32_t done = 0;
        while (!done)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(done) : "r"(smem_u32(bars + s)), "r"(parity) : "memory");
        const unsigned char *src = ring + s * kStageBytes;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int row = 8 * u + q;
            const float4 a = *reinterpret_cast<const float4 *>(src + row * 128 + 32 * t);
            const float4 c4 = *reinterpret_cast<const float4 *>(src + row * 128 + 32 * t + 16);
            const uint2 sv = *reinterpret_cast<const uint2 *>(src + 32 * 128 + row * 32 + 8 * t);
            acc += a.x + a.y + a.z + a.w + c4.x + c4.y + c4.z + c4.w + (float)(sv.x ^ sv.y);
        }
        __syncwarp();
    }
    if (acc == 12345.f) out[0] = acc;
}

template <typename F> static float time_ms(F f)
{
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    f();
    cudaDeviceSynchronize();
    cudaEventRecord(a);
    f();
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    return ms;
}

int main()
{
    float *out;
    cudaMalloc(&out, 4);
    const size_t nrows = 232965, nidx = 48u << 20;
    float *vals;
    unsigned char *sel;
    int *idx;
    cudaMalloc(&vals, nrows * 128);
    cudaMalloc(&sel, nrows * 32);
    cudaMalloc(&idx, nidx * 4);
    cudaMemset(vals, 0, nrows * 128);
    cudaMemset(sel, 0, nrows * 32);
    int *h = (int *)malloc(nidx * 4);
    unsigned long long s = 88172645463325252ull;
    for (size_t i = 0; i < nidx; ++i) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; h[i] = (int)(s % nrows); }
    cudaMemcpy(idx, h, nidx * 4, cudaMemcpyHostToDevice);
    free(h);
    for (int ctas : {2, 3, 4}) {
        const float ms = time_ms([&] { gather_ldg<<<148 * ctas, 256>>>(vals, sel, idx, nidx, out); });
        printf("ldg  4 lanes x (32 B + 8 B) per row, 8 rows/instr, %d CTAs/SM: %.3f ms, %.1f G rows/s\n", ctas, ms, nidx / ms / 1e6);
    }
    {
        constexpr int ST = 2;
        const int smem = 8 * (ST * 32 * 160 + 64);
        cudaFuncSetAttribute(gather_bulk<ST>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        for (int ctas : {1, 2}) {
            const float ms = time_ms([&] { gather_bulk<ST><<<148 * ctas, 256, smem>>>(vals, sel, idx, nidx, out); });
            printf("bulk cp.async.bulk 128 B + 32 B per row, %d stages, %d CTAs/SM: %.3f ms, %.1f G rows/s (%s)\n", ST, ctas, ms,
                   nidx / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
        }
    }
    {
        constexpr int ST = 4;
        const int smem = 8 * (ST * 32 * 160 + 64);
        cudaFuncSetAttribute(gather_bulk<ST>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        for (int ctas : {1}) {
            const float ms = time_ms([&] { gather_bulk<ST><<<148 * ctas, 256, smem>>>(vals, sel, idx, nidx, out); });
            printf("bulk cp.async.bulk 128 B + 32 B per row, %d stages, %d CTAs/SM: %.3f ms, %.1f G rows/s (%s)\n", ST, ctas, ms,
                   nidx / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
        }
    }
    return 0;
}
