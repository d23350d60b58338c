// Microbenchmark (developer tool): cost of one warp-wide "count keys >= t" step, the inner loop of the
// top-k search, in several formulations that load the ALU and FMA pipes differently.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_bin/count_bench tools/count_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int V>
__device__ __forceinline__ int lane_count(const uint32_t (&key)[8], const int (&skey)[8], uint32_t t, int one)
{
    if (V == 0) {   // borrow chain: 2 ALU-pipe IADD3 per key
        int c = 8;
        asm("{\n\t.reg .u32 d;\n\t"
            "sub.cc.u32 d, %1, %9;\n\tsubc.s32 %0, %0, 0;\n\t" "sub.cc.u32 d, %2, %9;\n\tsubc.s32 %0, %0, 0;\n\t"
            "sub.cc.u32 d, %3, %9;\n\tsubc.s32 %0, %0, 0;\n\t" "sub.cc.u32 d, %4, %9;\n\tsubc.s32 %0, %0, 0;\n\t"
            "sub.cc.u32 d, %5, %9;\n\tsubc.s32 %0, %0, 0;\n\t" "sub.cc.u32 d, %6, %9;\n\tsubc.s32 %0, %0, 0;\n\t"
            "sub.cc.u32 d, %7, %9;\n\tsubc.s32 %0, %0, 0;\n\t" "sub.cc.u32 d, %8, %9;\n\tsubc.s32 %0, %0, 0;\n\t}"
            : "+r"(c) : "r"(key[0]), "r"(key[1]), "r"(key[2]), "r"(key[3]), "r"(key[4]), "r"(key[5]), "r"(key[6]), "r"(key[7]), "r"(t));
        return c;
    } else if (V == 1) {   // compare + select-add (what the compiler emits for c += key >= t)
        int c = 0;
#pragma unroll
        for (int s = 0; s < 8; ++s) c += key[s] >= t ? 1 : 0;
        return c;
    } else if (V == 2) {   // signed difference + IMAD.HI sign accumulate: FMA pipe only (needs |key - t| < 2^31)
        const int st = (int)(t ^ 0x80000000u);
        int c = 8;
#pragma unroll
        for (int s = 0; s < 8; ++s) {
            int d = skey[s] - st;
            asm("mad.hi.s32 %0, %1, %2, %0;" : "+r"(c) : "r"(d), "r"(one));
        }
        return c;
    } else if (V == 3) {   // half the keys on each pipe
        const int st = (int)(t ^ 0x80000000u);
        int c = 8;
        asm("{\n\t.reg .u32 d;\n\t"
            "sub.cc.u32 d, %1, %5;\n\tsubc.s32 %0, %0, 0;\n\t" "sub.cc.u32 d, %2, %5;\n\tsubc.s32 %0, %0, 0;\n\t"
            "sub.cc.u32 d, %3, %5;\n\tsubc.s32 %0, %0, 0;\n\t" "sub.cc.u32 d, %4, %5;\n\tsubc.s32 %0, %0, 0;\n\t}"
            : "+r"(c) : "r"(key[0]), "r"(key[1]), "r"(key[2]), "r"(key[3]), "r"(t));
        int c2 = 0;
#pragma unroll
        for (int s = 4; s < 8; ++s) {
            int d = skey[s] - st;
            asm("mad.hi.s32 %0, %1, %2, %0;" : "+r"(c2) : "r"(d), "r"(one));
        }
        return c + c2;
    } else if (V == 5) {   // compare on the ALU pipe + predicated IMAD increment on the FMA pipe
        int c = 0;
        asm("{\n\t.reg .pred p;\n\t"
            "setp.ge.u32 p, %1, %9;\n\t@p mad.lo.s32 %0, %10, %10, %0;\n\t" "setp.ge.u32 p, %2, %9;\n\t@p mad.lo.s32 %0, %10, %10, %0;\n\t"
            "setp.ge.u32 p, %3, %9;\n\t@p mad.lo.s32 %0, %10, %10, %0;\n\t" "setp.ge.u32 p, %4, %9;\n\t@p mad.lo.s32 %0, %10, %10, %0;\n\t"
            "setp.ge.u32 p, %5, %9;\n\t@p mad.lo.s32 %0, %10, %10, %0;\n\t" "setp.ge.u32 p, %6, %9;\n\t@p mad.lo.s32 %0, %10, %10, %0;\n\t"
            "setp.ge.u32 p, %7, %9;\n\t@p mad.lo.s32 %0, %10, %10, %0;\n\t" "setp.ge.u32 p, %8, %9;\n\t@p mad.lo.s32 %0, %10, %10, %0;\n\t}"
            : "+r"(c) : "r"(key[0]), "r"(key[1]), "r"(key[2]), "r"(key[3]), "r"(key[4]), "r"(key[5]), "r"(key[6]), "r"(key[7]), "r"(t), "r"(one));
        return c;
    } else if (V == 6) {   // same with two accumulators (shorter dependency chain)
        int c = 0, c2 = 0;
        asm("{\n\t.reg .pred p, q;\n\t"
            "setp.ge.u32 p, %2, %10;\n\tsetp.ge.u32 q, %3, %10;\n\t@p mad.lo.s32 %0, %11, %11, %0;\n\t@q mad.lo.s32 %1, %11, %11, %1;\n\t"
            "setp.ge.u32 p, %4, %10;\n\tsetp.ge.u32 q, %5, %10;\n\t@p mad.lo.s32 %0, %11, %11, %0;\n\t@q mad.lo.s32 %1, %11, %11, %1;\n\t"
            "setp.ge.u32 p, %6, %10;\n\tsetp.ge.u32 q, %7, %10;\n\t@p mad.lo.s32 %0, %11, %11, %0;\n\t@q mad.lo.s32 %1, %11, %11, %1;\n\t"
            "setp.ge.u32 p, %8, %10;\n\tsetp.ge.u32 q, %9, %10;\n\t@p mad.lo.s32 %0, %11, %11, %0;\n\t@q mad.lo.s32 %1, %11, %11, %1;\n\t}"
            : "+r"(c), "+r"(c2) : "r"(key[0]), "r"(key[1]), "r"(key[2]), "r"(key[3]), "r"(key[4]), "r"(key[5]), "r"(key[6]), "r"(key[7]), "r"(t), "r"(one));
        return c + c2;
    } else {               // V == 4: difference via mad.lo (forced onto the FMA pipe) + IMAD.HI
        const int nst = -(int)(t ^ 0x80000000u);
        int c = 8;
#pragma unroll
        for (int s = 0; s < 8; ++s) {
            int d;
            asm("mad.lo.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(skey[s]), "r"(one), "r"(nst));
            asm("mad.hi.s32 %0, %1, %2, %0;" : "+r"(c) : "r"(d), "r"(one));
        }
        return c;
    }
}

template <int V>
__global__ void __launch_bounds__(256) bench(const uint32_t *keys, int iters, int one, uint32_t *out)
{
    uint32_t key[8];
    int skey[8];
    const int gid = blockIdx.x * blockDim.x + threadIdx.x;
#pragma unroll
    for (int s = 0; s < 8; ++s) {
        key[s] = keys[(gid * 8 + s) & 0xffff];
        skey[s] = (int)(key[s] ^ 0x80000000u);
    }
    uint32_t t = 0x80000000u + gid;
    uint32_t acc = 0;
    for (int i = 0; i < iters; ++i) {
        const int c = __reduce_add_sync(0xffffffffu, lane_count<V>(key, skey, t, one));
        acc += c;
        t = t * 1664525u + 1013904223u + c;     // next pivot depends on the count, like the search
        t = 0x40000000u + (t >> 1);            // keep |key - t| < 2^31 for keys in [0x40000000, 0xc0000000)
    }
    out[gid] = acc;
}

template <int V>
static void run(const char *name, const uint32_t *keys, uint32_t *out, uint32_t *ref)
{
    const int iters = 2000, grid = 148 * 4;
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    bench<V><<<grid, 256>>>(keys, 10, 1, out);
    cudaDeviceSynchronize();
    cudaEventRecord(a);
    bench<V><<<grid, 256>>>(keys, iters, 1, out);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms = 0;
    cudaEventElapsedTime(&ms, a, b);
    uint32_t h[256];
    cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
    bool same = true;
    if (V == 0) for (int i = 0; i < 256; ++i) ref[i] = h[i];
    else for (int i = 0; i < 256; ++i) same = same && ref[i] == h[i];
    // per SM sub-partition: 8 warps, iters counts each
    const double cyc = ms * 1e-3 * 1.93e9 / (iters * 8.0);
    printf("%-34s %.3f ms  %.1f cycles per count per scheduler  results %s\n", name, ms, cyc, same ? "match" : "DIFFER");
}

int main()
{
    uint32_t *keys, *out, h[65536], ref[256];
    for (int i = 0; i < 65536; ++i) h[i] = 0x40000000u + (uint32_t)((i * 2654435761u) >> 1);
    cudaMalloc(&keys, sizeof(h));
    cudaMalloc(&out, 148 * 4 * 256 * 4);
    cudaMemcpy(keys, h, sizeof(h), cudaMemcpyHostToDevice);
    run<0>("borrow chain (ALU x16)", keys, out, ref);
    run<1>("compare + select-add (compiler)", keys, out, ref);
    run<2>("sub + IMAD.HI (compiler's choice)", keys, out, ref);
    run<3>("4 borrow + 4 IMAD.HI", keys, out, ref);
    run<4>("IMAD + IMAD.HI (FMA x16)", keys, out, ref);
    run<5>("ISETP + predicated IMAD", keys, out, ref);
    run<6>("ISETP + predicated IMAD, 2 chains", keys, out, ref);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
