// Developer microbenchmark: issue cost of cp.async.bulk (UBLKCP), mbarrier try_wait, cp.async+arrive.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned s32(const void* p){ return (unsigned)__cvta_generic_to_shared(p); }
__global__ void k(const float* g, long long* out, int bytes, int ncopies){
  extern __shared__ __align__(128) unsigned char sm[];
  uint64_t* bar=(uint64_t*)sm; float* dst=(float*)(sm+128);
  if(threadIdx.x==0){ asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;"::"r"(s32(bar))); asm volatile("fence.mbarrier_init.release.cluster;"); asm volatile("fence.proxy.async;"); }
  __syncthreads();
  long long t_issue=0,t_wait=0,t_wait2=0; unsigned par=0;
  for(int it=0; it<64; ++it){
    if(threadIdx.x==0){
      long long t0=clock64();
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;"::"r"(s32(bar)),"r"(bytes*ncopies):"memory");
      for(int c=0;c<ncopies;++c)
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"::"r"(s32(dst+c*(bytes/4))),"l"(g+(size_t)(it*ncopies+c)*(bytes/4)),"r"(bytes),"r"(s32(bar)):"memory");
      long long t1=clock64();
      unsigned ok=0; while(!ok){ asm volatile("{.reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0,1,0,p;}":"=r"(ok):"r"(s32(bar)),"r"(par):"memory"); }
      long long t2=clock64();
      ok=0; while(!ok){ asm volatile("{.reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0,1,0,p;}":"=r"(ok):"r"(s32(bar)),"r"(par):"memory"); }
      long long t3=clock64();
      par^=1; t_issue+=t1-t0; t_wait+=t2-t1; t_wait2+=t3-t2;
    }
    __syncthreads();
  }
  if(threadIdx.x==0){ out[0]=t_issue/64; out[1]=t_wait/64; out[2]=t_wait2/64; }
}
int main(){
  float* g; cudaMalloc(&g, 64<<20); cudaMemset(g,0,64<<20); long long* o; cudaMalloc(&o,64);
  int cfgs[][2]={{192,1},{192,4},{2064,1},{2064,4},{8256,1},{16512,1}};
  for(auto& c:cfgs){
    k<<<1,32,128+70000>>>(g,o,c[0],c[1]); cudaDeviceSynchronize();
    cudaFuncSetAttribute(k,cudaFuncAttributeMaxDynamicSharedMemorySize,128+70000);
    k<<<1,32,128+70000>>>(g,o,c[0],c[1]); cudaDeviceSynchronize();
    long long h[3]; cudaMemcpy(h,o,24,cudaMemcpyDeviceToHost);
    printf("bytes=%5d x%d: issue %lld cyc, wait-until-landed %lld cyc, try_wait on completed phase %lld cyc (%s)\n",c[0],c[1],h[0],h[1],h[2],cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
