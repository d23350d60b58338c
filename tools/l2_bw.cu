// Microbenchmarks (developer tool): L2 read bandwidth and random CBSR-row gather rate on B200.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
__global__ void rd(const float4* __restrict__ p, size_t n4, int iters, float* out){
  float4 acc = make_float4(0,0,0,0);
  size_t stride = (size_t)gridDim.x*blockDim.x;
  for(int it=0; it<iters; ++it){
    size_t tid = ((blockIdx.x + it*37) % gridDim.x)*(size_t)blockDim.x+threadIdx.x;   // rotate slices so L1 never helps
    for(size_t i=tid;i<n4;i+=stride){ float4 v=__ldg(p+i); acc.x+=v.x; acc.y+=v.y; acc.z+=v.z; acc.w+=v.w; }
  }
  if(acc.x+acc.y+acc.z+acc.w==12345.f) out[0]=acc.x;
}
// mode 0: lane-per-4B, one row (128 B) per instruction, optional 32 B selector row
// mode 1: 8 lanes x 16 B per row, 4 rows per instruction (vals), selectors 8 lanes x 4 B
// mode 2: packed 160 B records, 10 lanes x 16 B, 3 rows per instruction
template<int MODE, int U>
__global__ void gather(const float* __restrict__ vals, const unsigned char* __restrict__ sel, const int* __restrict__ idx, size_t nidx, int with_sel, float* out){
  size_t warp = (blockIdx.x*(size_t)blockDim.x+threadIdx.x)>>5, nwarps=((size_t)gridDim.x*blockDim.x)>>5; int lane=threadIdx.x&31;
  float acc=0;
  if (MODE==0){
    for(size_t i=warp*U;i+U<=nidx;i+=nwarps*U){
      float v[U]; int s[U];
      #pragma unroll
      for(int u=0;u<U;++u){ int c=__ldg(idx+i+u); v[u]=__ldg(vals+(size_t)c*32+lane); s[u]= with_sel? __ldg(sel+(size_t)c*32+lane):0; }
      #pragma unroll
      for(int u=0;u<U;++u) acc+=v[u]+s[u];
    }
  } else if (MODE==1){
    const int q=lane>>3, t=lane&7;
    for(size_t i=warp*U*4;i+U*4<=nidx;i+=nwarps*U*4){
      float4 v[U]; unsigned s[U];
      #pragma unroll
      for(int u=0;u<U;++u){ int c=__ldg(idx+i+u*4+q); v[u]=__ldg(reinterpret_cast<const float4*>(vals+(size_t)c*32)+t); s[u]= with_sel? __ldg(reinterpret_cast<const unsigned*>(sel+(size_t)c*32)+t):0u; }
      #pragma unroll
      for(int u=0;u<U;++u) acc+=v[u].x+v[u].y+v[u].z+v[u].w+s[u];
    }
  } else {
    const int q=lane/10, t=lane%10;   // lanes 30,31 idle
    for(size_t i=warp*U*3;i+U*3<=nidx;i+=nwarps*U*3){
      float4 v[U];
      #pragma unroll
      for(int u=0;u<U;++u){ int c=__ldg(idx+i+u*3+(q<3?q:0)); v[u]= q<3 ? __ldg(reinterpret_cast<const float4*>(vals+(size_t)c*40)+t) : make_float4(0,0,0,0); }
      #pragma unroll
      for(int u=0;u<U;++u) acc+=v[u].x+v[u].y+v[u].z+v[u].w;
    }
  }
  if(acc==12345.f) out[0]=acc;
}
// RED test: each warp adds 128-byte rows (8 lanes x red.v4) to random rows, 4 rows per instruction
__global__ void redtest(float* __restrict__ dst, const int* __restrict__ idx, size_t nidx, int vec){
  size_t warp = (blockIdx.x*(size_t)blockDim.x+threadIdx.x)>>5, nwarps=((size_t)gridDim.x*blockDim.x)>>5; int lane=threadIdx.x&31;
  const int q=lane>>3, t=lane&7;
  for(size_t i=warp*16;i+16<=nidx;i+=nwarps*16){
    #pragma unroll
    for(int u=0;u<4;++u){
      if(vec){ int c=__ldg(idx+i+u*4+q); float* p=dst+(size_t)c*32+4*t;
        asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" :: "l"(p), "f"(1.f),"f"(1.f),"f"(1.f),"f"(1.f) : "memory"); }
      else { for(int v=0;v<4;++v){ int c=__ldg(idx+i+u*4+v); atomicAdd(dst+(size_t)c*32+lane, 1.f);} }
    }
  }
}
template<int MODE,int U>
void run_gather(const char* name, float* vals, unsigned char* sel, int* idx, size_t nidx, size_t nrows, int ws, float* out, int ctas, int threads){
  cudaEvent_t a,b; cudaEventCreate(&a); cudaEventCreate(&b);
  gather<MODE,U><<<148*ctas,threads>>>(vals,sel,idx,nidx,ws,out); cudaDeviceSynchronize();
  cudaEventRecord(a); gather<MODE,U><<<148*ctas,threads>>>(vals,sel,idx,nidx,ws,out); cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms,a,b);
  printf("gather %-28s rows=%zu sel=%d ctas/SM=%d thr=%d U=%d: %.3f ms, %.1f G rows/s, %.1f GB/s\n", name, nrows, ws, ctas, threads, U, ms, nidx/ms/1e6, nidx*(128.0+32*ws)/ms/1e6);
}
int main(){
  float* out; cudaMalloc(&out,4);
  cudaEvent_t a,b; cudaEventCreate(&a); cudaEventCreate(&b);
  for(size_t mb : {16,32,48,64,96,256,1024}){
    size_t bytes=mb<<20; float4* p; cudaMalloc(&p,bytes); cudaMemset(p,0,bytes);
    int iters = mb<=96?20:4;
    rd<<<148*8,512>>>(p,bytes/16,2,out); cudaDeviceSynchronize();
    cudaEventRecord(a); rd<<<148*8,512>>>(p,bytes/16,iters,out); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms,a,b);
    printf("read %4zu MB x%d: %.1f GB/s\n", mb, iters, bytes*(double)iters/ms/1e6);
    cudaFree(p);
  }
  for(size_t nrows : {232965ul, 2449029ul}){
    float* vals; unsigned char* sel; int* idx; size_t nidx = 48u<<20;
    cudaMalloc(&vals,nrows*160); cudaMalloc(&sel,nrows*32); cudaMalloc(&idx,nidx*4);
    cudaMemset(vals,0,nrows*160); cudaMemset(sel,0,nrows*32);
    int* h=(int*)malloc(nidx*4); unsigned long long s=88172645463325252ull;
    for(size_t i=0;i<nidx;++i){ s^=s<<13; s^=s>>7; s^=s<<17; h[i]=(int)(s%nrows);}
    cudaMemcpy(idx,h,nidx*4,cudaMemcpyHostToDevice); free(h);
    run_gather<0,8>("lane4B 1row/instr", vals,sel,idx,nidx,nrows,0,out,8,256);
    run_gather<0,8>("lane4B 1row/instr", vals,sel,idx,nidx,nrows,1,out,8,256);
    run_gather<0,4>("lane4B 1row/instr", vals,sel,idx,nidx,nrows,1,out,4,256);
    run_gather<0,16>("lane4B 1row/instr", vals,sel,idx,nidx,nrows,1,out,4,256);
    run_gather<1,4>("8lanes16B 4rows/instr", vals,sel,idx,nidx,nrows,0,out,8,256);
    run_gather<1,4>("8lanes16B 4rows/instr", vals,sel,idx,nidx,nrows,1,out,8,256);
    run_gather<1,2>("8lanes16B 4rows/instr", vals,sel,idx,nidx,nrows,1,out,8,256);
    run_gather<1,8>("8lanes16B 4rows/instr", vals,sel,idx,nidx,nrows,1,out,4,256);
    run_gather<2,4>("packed160B 3rows/instr", vals,sel,idx,nidx,nrows,1,out,8,256);
    for(int vec=0; vec<2; ++vec){
      cudaEvent_t a2,b2; cudaEventCreate(&a2); cudaEventCreate(&b2);
      redtest<<<148*8,256>>>(vals,idx,nidx,vec); cudaDeviceSynchronize();
      cudaEventRecord(a2); redtest<<<148*8,256>>>(vals,idx,nidx,vec); cudaEventRecord(b2); cudaEventSynchronize(b2);
      float ms; cudaEventElapsedTime(&ms,a2,b2);
      printf("RED 128B rows=%zu vec4=%d: %.3f ms, %.1f G rows/s, %.1f GB/s\n", nrows, vec, ms, nidx/ms/1e6, nidx*128.0/ms/1e6);
    }
    cudaFree(vals); cudaFree(sel); cudaFree(idx);
  }
  return 0;
}
