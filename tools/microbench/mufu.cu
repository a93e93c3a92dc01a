// Developer microbenchmark: MUFU / FADD / SHFL / LDS latency and throughput per warp on sm_100a.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ float ex2f(float x){float y; asm volatile("ex2.approx.ftz.f32 %0, %1;":"=f"(y):"f"(x)); return y;}
__device__ __forceinline__ float lg2f(float x){float y; asm volatile("lg2.approx.ftz.f32 %0, %1;":"=f"(y):"f"(x)); return y;}
template<int MODE, int ILP>
__global__ void k(float* out, long long* cyc, int iters){
  float v[ILP];
  for(int i=0;i<ILP;++i) v[i]=0.5f+0.001f*threadIdx.x+i;
  __shared__ float sm[1024];
  sm[threadIdx.x]=threadIdx.x;
  __syncthreads();
  long long t0=clock64();
  for(int it=0;it<iters;++it){
#pragma unroll
    for(int i=0;i<ILP;++i){
      if(MODE==0) v[i]=ex2f(v[i]);                 // dependent per chain
      if(MODE==1) v[i]=lg2f(v[i]);
      if(MODE==2) v[i]=v[i]+1.0f;
      if(MODE==3) v[i]=__shfl_up_sync(0xffffffffu,v[i],1);
      if(MODE==4) v[i]=sm[((int)v[i])&1023];
      if(MODE==5) { float a=v[i]; v[i]=fmaxf(a,0.25f)+lg2f(1.0f+ex2f(-fabsf(a-0.25f))); }  // lse2
      if(MODE==6) { float m; asm volatile("redux.sync.max.f32 %0, %1, 0xffffffff;" : "=f"(m) : "f"(v[i])); v[i]=m+1.0f; }
      if(MODE==7) { float x=v[i]; for(int o=16;o>0;o>>=1) x+=__shfl_xor_sync(0xffffffffu,x,o); v[i]=x*0.03f; }
      if(MODE==8) { float x=v[i]; for(int o=16;o>0;o>>=1) x=fmaxf(x,__shfl_xor_sync(0xffffffffu,x,o)); v[i]=x+1.0f; }
    }
  }
  long long t1=clock64();
  float s=0; for(int i=0;i<ILP;++i) s+=v[i];
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
  if(threadIdx.x==0) cyc[blockIdx.x]=t1-t0;
}
template<int MODE,int ILP> void run(const char* name,int warps){
  float* out; long long* cyc; cudaMalloc(&out,4096*4); cudaMalloc(&cyc,8*8);
  int iters=2000;
  k<MODE,ILP><<<1,32*warps>>>(out,cyc,iters); cudaDeviceSynchronize();
  k<MODE,ILP><<<1,32*warps>>>(out,cyc,iters); cudaDeviceSynchronize();
  long long h; cudaMemcpy(&h,cyc,8,cudaMemcpyDeviceToHost);
  printf("%-8s ILP=%d warps=%d : %.2f cycles per op-group (=%.2f per op)\n",name,ILP,warps,(double)h/iters,(double)h/iters/ILP);
  cudaFree(out); cudaFree(cyc);
}
int main(){
  run<0,1>("ex2",1); run<0,4>("ex2",1); run<0,16>("ex2",1); run<0,16>("ex2",4); run<0,16>("ex2",8);
  run<1,1>("lg2",1); run<1,16>("lg2",1); run<1,16>("lg2",4);
  run<2,1>("fadd",1); run<2,16>("fadd",1); run<2,16>("fadd",4); run<2,16>("fadd",8);
  run<3,1>("shfl",1); run<3,8>("shfl",1);
  run<4,1>("lds",1); run<4,8>("lds",1);
  run<6,1>("credux",1); run<6,4>("credux",1); run<7,1>("shflsum",1); run<7,4>("shflsum",1); run<8,1>("shflmax",1);
  run<5,1>("lse2",1); run<5,4>("lse2",1); run<5,8>("lse2",1); run<5,8>("lse2",4); run<5,8>("lse2",8);
  return 0;
}
