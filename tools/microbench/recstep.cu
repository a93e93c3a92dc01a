// Developer microbenchmark: the lattice recursion step (P pairs per thread) in isolation.
#include <cstdio>
#include <cuda_runtime.h>
#include "../../pytorch-asr_b200/csrc/ctc_kernels.cuh"
using namespace ctcb200;
template<int P, int STORE, int NEGINIT>
__global__ void k(float* lat, long long* cyc, int steps, int V, int RS){
  __shared__ float lp2[8*64];
  const int lane=threadIdx.x&31, w=threadIdx.x>>5, tid=threadIdx.x;
  for(int i=tid;i<8*64;i+=blockDim.x) lp2[i]=-3.f-0.01f*(i%37);
  __syncthreads();
  float aB[P], aY[P]; int lab[P]; bool skip[P];
  for(int k2=0;k2<P;++k2){aB[k2]=NEGINIT? kNeg : -1.f*tid; aY[k2]=NEGINIT? kNeg : -2.f*tid; if(NEGINIT && tid==0 && k2==0) aB[0]=0.f; lab[k2]=(tid*7+k2*3)%V; skip[k2]=((tid+k2)%5)!=0;}
  const int i0=tid*P; float off=0.f;
  long long t0=clock64();
  for(int tt=0;tt<steps;++tt){
    const float* row=lp2+(tt&7)*64;
    const float lpb=row[0];
    float am1=__shfl_up_sync(0xffffffffu,aY[P-1],1);
    if(lane==0) am1=-1e30f;
    float lpl[P];
#pragma unroll
    for(int k2=0;k2<P;++k2){
      lpl[k2]=row[lab[k2]];
      const float x=lse2(aB[k2],am1);
      const float yin=skip[k2]?x:aB[k2];
      const float ynew=lpl[k2]+lse2(aY[k2],yin);
      am1=aY[k2]; aB[k2]=lpb+x; aY[k2]=ynew;
    }
    if((tt&7)==7){
      float m=kNeg;
#pragma unroll
      for(int k2=0;k2<P;++k2) m=fmaxf(m,fmaxf(aB[k2],aY[k2]));
      m=warp_max(m);
      const float sh=(m>kRealThresh)?rintf(m):0.f;
#pragma unroll
      for(int k2=0;k2<P;++k2){aB[k2]=fmaxf(aB[k2]-sh,kNeg); aY[k2]=fmaxf(aY[k2]-sh,kNeg);}
      off+=sh;
    }
    if(STORE){
      float* r=lat+(size_t)tt*RS;
      store_vec<P>(r+i0,aB); store_vec<P>(r+RS/2+i0,aY);
    }
  }
  long long t1=clock64();
  float s=off; for(int k2=0;k2<P;++k2) s+=aB[k2]+aY[k2];
  lat[tid]=s;
  if(tid==0) cyc[0]=t1-t0;
}
template<int P,int STORE,int NEGINIT> void run(int warps){
  float* lat; long long* cyc; int steps=1000, RS=2*32*P*warps;
  cudaMalloc(&lat,(size_t)steps*RS*4+4096); cudaMalloc(&cyc,8);
  k<P,STORE,NEGINIT><<<1,32*warps>>>(lat,cyc,steps,48,RS); cudaDeviceSynchronize();
  k<P,STORE,NEGINIT><<<1,32*warps>>>(lat,cyc,steps,48,RS); cudaDeviceSynchronize();
  long long h; cudaMemcpy(&h,cyc,8,cudaMemcpyDeviceToHost);
  printf("P=%d store=%d neginit=%d warps=%d: %.1f cycles/step  (%s)\n",P,STORE,NEGINIT,warps,(double)h/steps,cudaGetErrorString(cudaGetLastError()));
  cudaFree(lat); cudaFree(cyc);
}
int main(){ run<4,0,0>(1); run<4,1,0>(1); run<4,0,1>(1); run<4,1,1>(1); run<2,1,1>(1); run<1,1,1>(1); return 0; }
