// Microbenchmark (developer tool): 128-byte reductions into random rows of an L2-resident / HBM-resident
// fp32 matrix, (a) red.global.add.v4.f32 from registers, (b) cp.reduce.async.bulk (TMA reduce) from shared memory.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>

__global__ void red_v4(float* __restrict__ dst, const int* __restrict__ idx, size_t nidx){
  size_t warp = (blockIdx.x*(size_t)blockDim.x+threadIdx.x)>>5, nwarps=((size_t)gridDim.x*blockDim.x)>>5; int lane=threadIdx.x&31;
  const int q=lane>>3, t=lane&7;
  for(size_t i=warp*16;i+16<=nidx;i+=nwarps*16){
    #pragma unroll
    for(int u=0;u<4;++u){
      int c=__ldg(idx+i+u*4+q); float* p=dst+(size_t)c*32+4*t;
      asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" :: "l"(p), "f"(1.f),"f"(1.f),"f"(1.f),"f"(1.f) : "memory");
    }
  }
}
// each warp: stage 32 rows x 128 B in shared memory (values already there), one lane per row issues a bulk reduce
template<int ROWS_PER_ISSUE>
__global__ void red_tma(float* __restrict__ dst, const int* __restrict__ idx, size_t nidx){
  extern __shared__ __align__(128) float stage[];   // [warps][32 rows][32 floats]
  const int lane=threadIdx.x&31, w=threadIdx.x>>5;
  float* my = stage + (size_t)w*32*32;
  for(int i=lane;i<32*32;i+=32) my[i]=1.f;
  __syncwarp();
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  size_t warp = (blockIdx.x*(size_t)blockDim.x+threadIdx.x)>>5, nwarps=((size_t)gridDim.x*blockDim.x)>>5;
  for(size_t i=warp*32;i+32<=nidx;i+=nwarps*32){
    int c=__ldg(idx+i+lane);
    float* p=dst+(size_t)c*32;
    uint32_t s=(uint32_t)__cvta_generic_to_shared(my + lane*32);
    asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], 128;" :: "l"(p), "r"(s) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 4;" ::: "memory");
  }
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}
int main(){
  cudaEvent_t a,b; cudaEventCreate(&a); cudaEventCreate(&b);
  for(size_t nrows : {232965ul, 2449029ul}){
    float* dst; int* idx; size_t nidx = 48u<<20;
    cudaMalloc(&dst,nrows*128); cudaMalloc(&idx,nidx*4); cudaMemset(dst,0,nrows*128);
    int* h=(int*)malloc(nidx*4); unsigned long long s=88172645463325252ull;
    for(size_t i=0;i<nidx;++i){ s^=s<<13; s^=s>>7; s^=s<<17; h[i]=(int)(s%nrows);}
    cudaMemcpy(idx,h,nidx*4,cudaMemcpyHostToDevice); free(h);
    float ms;
    red_v4<<<148*8,256>>>(dst,idx,nidx); cudaDeviceSynchronize();
    cudaEventRecord(a); red_v4<<<148*8,256>>>(dst,idx,nidx); cudaEventRecord(b); cudaEventSynchronize(b); cudaEventElapsedTime(&ms,a,b);
    printf("rows=%zu red.v4 : %.3f ms  %.1f G rows/s\n", nrows, ms, nidx/ms/1e6);
    cudaFuncSetAttribute(red_tma<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8*32*32*4);
    red_tma<1><<<148*4,256,8*32*32*4>>>(dst,idx,nidx); cudaError_t e=cudaDeviceSynchronize();
    if(e!=cudaSuccess){ printf("tma reduce failed: %s\n", cudaGetErrorString(e)); return 1; }
    cudaEventRecord(a); red_tma<1><<<148*4,256,8*32*32*4>>>(dst,idx,nidx); cudaEventRecord(b); cudaEventSynchronize(b); cudaEventElapsedTime(&ms,a,b);
    printf("rows=%zu cp.reduce.async.bulk 128B : %.3f ms  %.1f G rows/s\n", nrows, ms, nidx/ms/1e6);
    // check: total sum must equal 2 passes x (2 kernels) x nidx rows x 32 ... just sample
    float hv[32]; cudaMemcpy(hv,dst,128,cudaMemcpyDeviceToHost); printf("  sample dst[0][0..1] = %.0f %.0f\n", hv[0], hv[1]);
    cudaFree(dst); cudaFree(idx);
  }
  return 0;
}
