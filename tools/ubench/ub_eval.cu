// Micro-benchmark: the region-1 evaluation sequence of k_voigt_tile (10 FP64 ops + MUFU.RCP64H per
// eval) with parameters in registers, for different ILP / warps per SM.  Gives the ceiling the
// tile kernel can reach with its instruction mix.  nvcc -O3 -gencode arch=compute_100a,code=sm_100a -cudart shared
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ double rcp_approx(double x){double r; asm("rcp.approx.ftz.f64 %0, %1;":"=d"(r):"d"(x)); return r;}
__device__ __forceinline__ double reg1_fast(double u,double c2){double den=fma(u,u,c2);double w=u+1.0;double r0=rcp_approx(den);double e=fma(-den,r0,1.0);double t=w*r0;return fma(t,e,t);}
template<int ILP,bool DIV>
__global__ void __launch_bounds__(512) ub(int iters,double A0,double B,double C,double c2,double* sink){
  double pd[ILP],a0[ILP],a1[ILP],a2[ILP];
  #pragma unroll
  for(int k=0;k<ILP;k++){pd[k]=threadIdx.x+k*blockDim.x;a0[k]=a1[k]=a2[k]=0;}
  double A=A0;
  for(int i=0;i<iters;i++){
    #pragma unroll
    for(int k=0;k<ILP;k++){
      double u=fma(fma(C,pd[k],B),pd[k],A);
      double kp;
      if(DIV){ double den=fma(u,u,c2); kp=(u+1.0)/den; } else kp=reg1_fast(u,c2);
      a0[k]=fma(1.1,kp,a0[k]);a1[k]=fma(1.2,kp,a1[k]);a2[k]=fma(1.3,kp,a2[k]);
    }
    A+=1e-3;
  }
  double s=0;
  #pragma unroll
  for(int k=0;k<ILP;k++)s+=a0[k]+a1[k]+a2[k];
  if(s==1.2345)sink[0]=s;
}
template<int ILP,bool DIV> void run(int blocks_per_sm,int threads,int iters){
  double* sink;cudaMalloc(&sink,8);
  cudaEvent_t e0,e1;cudaEventCreate(&e0);cudaEventCreate(&e1);
  int blocks=148*blocks_per_sm;
  ub<ILP,DIV><<<blocks,threads>>>(iters/4,300.0,0.5,1e-3,1e-9,sink);
  cudaEventRecord(e0);
  ub<ILP,DIV><<<blocks,threads>>>(iters,300.0,0.5,1e-3,1e-9,sink);
  cudaEventRecord(e1);cudaEventSynchronize(e1);
  float ms;cudaEventElapsedTime(&ms,e0,e1);
  double evals=(double)iters*ILP*blocks*threads;
  printf("ILP %d div %d blocks/SM %d threads %d (warps/SM %d): %.3e evals/s\n",ILP,(int)DIV,blocks_per_sm,threads,blocks_per_sm*threads/32,evals/(ms*1e-3));
  cudaFree(sink);
}
int main(){
  int it=20000;
  run<1,false>(1,128,it); run<2,false>(1,128,it); run<4,false>(1,128,it); run<8,false>(1,128,it);
  run<1,false>(1,256,it); run<2,false>(1,256,it); run<4,false>(1,256,it); run<8,false>(1,256,it);
  run<1,false>(1,512,it); run<2,false>(1,512,it); run<4,false>(1,512,it);
  run<1,false>(2,512,it); run<2,false>(2,512,it); run<2,false>(4,512,it); run<4,false>(4,512,it);
  run<2,true>(1,256,it); run<4,true>(4,512,it);
  return 0;
}
