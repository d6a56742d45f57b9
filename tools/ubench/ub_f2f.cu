// Micro-benchmark: throughput of F2F.F64.F32 (float->double) and of the LUT inner loop variants
#include <cstdio>
#include <cuda_runtime.h>
template<int MODE>
__global__ void __launch_bounds__(256) ub(int iters, const float* __restrict__ src, double w, double* sink){
  // 8 floats per thread in registers, re-converted each iteration with a changing bit so the
  // compiler cannot hoist the conversion
  float f[8];
  #pragma unroll
  for(int k=0;k<8;k++) f[k]=src[threadIdx.x*8+k];
  double acc0=0,acc1=0; float facc0=0,facc1=0;
  for(int i=0;i<iters;i++){
    #pragma unroll
    for(int k=0;k<8;k++){
      float v=__int_as_float(__float_as_int(f[k])^(i&1));
      if(MODE==0){ double d=(double)v; if(k&1) acc1=fma(w,d,acc1); else acc0=fma(w,d,acc0);}          // F2F + DFMA
      else if(MODE==1){ if(k&1) facc1=fmaf((float)w,v,facc1); else facc0=fmaf((float)w,v,facc0);}     // FFMA only
      else if(MODE==2){ double d=(double)v; acc0+=d; }                                               // F2F + DADD
      else { double d=__hiloint2double(__float_as_int(v),0); if(k&1) acc1=fma(w,d,acc1); else acc0=fma(w,d,acc0);} // no F2F
    }
  }
  if(acc0+acc1+facc0+facc1==1.2345) sink[0]=acc0;
}
template<int MODE> void run(const char* name,float* src,double* sink){
  cudaEvent_t e0,e1;cudaEventCreate(&e0);cudaEventCreate(&e1);
  int it=20000; int blocks=148*8;
  ub<MODE><<<blocks,256>>>(it/4,src,1.0001,sink);
  cudaEventRecord(e0); ub<MODE><<<blocks,256>>>(it,src,1.0001,sink); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms;cudaEventElapsedTime(&ms,e0,e1);
  double ops=(double)it*8*blocks*256;
  printf("%s: %.3e elem/s = %.2f per clk per SM (at 1.9 GHz)\n",name,ops/(ms*1e-3),ops/(ms*1e-3)/148/1.9e9);
}
int main(){ float* src; double* sink; cudaMalloc(&src,256*8*4); cudaMemset(src,0x3f,256*8*4); cudaMalloc(&sink,8);
  run<0>("F2F+DFMA",src,sink); run<1>("FFMA",src,sink); run<2>("F2F+DADD",src,sink); run<3>("bits+DFMA (no F2F)",src,sink); return 0;}
