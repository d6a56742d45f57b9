// Micro-benchmark: the far-wing record loop of k_voigt_tile exactly as the kernel runs it - 64-byte
// HalfRec records broadcast from shared memory, PPT points per thread, two records per iteration -
// to separate the ceiling of THIS loop from the rest of the kernel (loader, centres, stores).
// nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -cudart shared -o ub_tile_loop ub_tile_loop.cu
#include <cstdio>
#include <cuda_runtime.h>
struct __align__(16) HalfRec { double xs, b, c1, c2, g0, g1, g2; int lo; unsigned len; };
__device__ __forceinline__ double rcp_approx(double x){double r; asm("rcp.approx.ftz.f64 %0, %1;":"=d"(r):"d"(x)); return r;}
__device__ __forceinline__ double reg1_uw(double u,double w,double c2){double den=fma(u,u,c2);double r0=rcp_approx(den);double e=fma(-den,r0,1.0);double t=w*r0;return fma(t,e,t);}

template<int NT,int PPT,int MINB,int VAR>
__global__ void __launch_bounds__(NT,MINB) ub(int n_rec,int reps,double* sink){
  extern __shared__ __align__(16) unsigned char sm[];
  HalfRec* hbuf=reinterpret_cast<HalfRec*>(sm);
  for(int i=threadIdx.x;i<n_rec;i+=NT){HalfRec h; h.xs=1e-3+1e-7*i; h.b=300.0+i; h.c1=0.25; h.c2=1e-9; h.g0=1.1; h.g1=1.2; h.g2=1.3; h.lo=0; h.len=1u<<30; hbuf[i]=h;}
  __syncthreads();
  double Pd[PPT],acc0[PPT],acc1[PPT],acc2[PPT];
  #pragma unroll
  for(int k=0;k<PPT;k++){Pd[k]=(double)(blockIdx.x*NT*PPT+threadIdx.x+k*NT); asm volatile("":"+d"(Pd[k])); acc0[k]=acc1[k]=acc2[k]=0.0;}
  auto half_eval=[&](const HalfRec& h){
    const double xs=h.xs,b=h.b,c1=h.c1,c2=h.c2,g0=h.g0,g1=h.g1,g2=h.g2;
    const double c1p=c1+1.0;
    #pragma unroll
    for(int k=0;k<PPT;k++){
      const double x=fma(Pd[k],xs,b);
      double kp=reg1_uw(fma(x,x,c1),fma(x,x,c1p),c2);
      acc0[k]=fma(g0,kp,acc0[k]);acc1[k]=fma(g1,kp,acc1[k]);acc2[k]=fma(g2,kp,acc2[k]);
    }
  };
  const int P0=blockIdx.x*NT*PPT+threadIdx.x;
  auto half_part=[&](const HalfRec& h, int mode){
    const double xs=h.xs,b=h.b,c1=h.c1,c2=h.c2,g0=h.g0,g1=h.g1,g2=h.g2;
    const double c1p=c1+1.0;
    const int lo=h.lo; const unsigned len=h.len;
    if(mode==2){ bool anyk=false;
      #pragma unroll
      for(int k=0;k<PPT;k++) anyk|=(unsigned)(P0+k*NT-lo)<=len;
      if(!__any_sync(0xffffffffu,anyk)) return; }
    #pragma unroll
    for(int k=0;k<PPT;k++){
      bool in=(unsigned)(P0+k*NT-lo)<=len;
      if(mode==0){ if(!__any_sync(0xffffffffu,in)) continue; }
      const double x=fma(Pd[k],xs,b);
      double kp=reg1_uw(fma(x,x,c1),fma(x,x,c1p),c2);
      kp=in?kp:0.0;
      acc0[k]=fma(g0,kp,acc0[k]);acc1[k]=fma(g1,kp,acc1[k]);acc2[k]=fma(g2,kp,acc2[k]);
    }
  };
  if(VAR>=10){ // partial records: window edge uniformly inside this CTA's tile
    for(int i=threadIdx.x;i<n_rec;i+=NT){ unsigned edge=(unsigned)((i*2654435761u)>>8)%(NT*PPT); bool left=i&1;
      hbuf[i].lo = left ? (int)(blockIdx.x*NT*PPT) - 100000 : (int)(blockIdx.x*NT*PPT+edge);
      hbuf[i].len= left ? (unsigned)(100000+edge) : 1u<<30; }
    __syncthreads();
  }
  // FP32 evaluation of K, FP64 accumulation (d = P - centre in float, then the same rational)
  float Pf[PPT];
  #pragma unroll
  for(int k=0;k<PPT;k++) Pf[k]=(float)(threadIdx.x+k*NT);
  auto half_f32=[&](const HalfRec& h){
    const float xs=(float)h.xs,b=(float)(h.b-300.0),c1=(float)h.c1,c2=(float)h.c2;
    const double g0=h.g0,g1=h.g1,g2=h.g2;
    const float c1p=c1+1.0f;
    #pragma unroll
    for(int k=0;k<PPT;k++){
      const float x=fmaf(Pf[k],xs,b);
      const float u=fmaf(x,x,c1), w=fmaf(x,x,c1p);
      const float den=fmaf(u,u,c2);
      float rr; asm("rcp.approx.ftz.f32 %0, %1;":"=f"(rr):"f"(den)); const double kp=(double)(w*rr);
      acc0[k]=fma(g0,kp,acc0[k]);acc1[k]=fma(g1,kp,acc1[k]);acc2[k]=fma(g2,kp,acc2[k]);
    }
  };
  if(VAR==20){ for(int r=0;r<reps;r++){ int h=0; for(;h+2<=n_rec;h+=2){half_f32(hbuf[h]);half_f32(hbuf[h+1]);} } }
  for(int r=0;r<reps && VAR!=20;r++){
    if(VAR>=10){ for(int h=0;h<n_rec;h++) half_part(hbuf[h],VAR-10); }
    else if(VAR==0){ int h=0; for(;h+2<=n_rec;h+=2){half_eval(hbuf[h]);half_eval(hbuf[h+1]);} }
    else if(VAR==1){ for(int h=0;h<n_rec;h++) half_eval(hbuf[h]); }
    else if(VAR==2){ // software-pipelined record fetch: next record in registers while this one computes
      HalfRec cur=hbuf[0];
      for(int h=0;h<n_rec;h++){ HalfRec nxt=hbuf[h+1<n_rec?h+1:h]; half_eval(cur); cur=nxt; }
    } else { int h=0; for(;h+4<=n_rec;h+=4){half_eval(hbuf[h]);half_eval(hbuf[h+1]);half_eval(hbuf[h+2]);half_eval(hbuf[h+3]);} }
  }
  double s=0;
  #pragma unroll
  for(int k=0;k<PPT;k++)s+=acc0[k]+acc1[k]+acc2[k];
  if(s==1.2345)sink[0]=s;
}
template<int NT,int PPT,int MINB,int VAR> void run(int blocks_per_sm,int n_rec,int reps){
  double* sink;cudaMalloc(&sink,8);
  cudaEvent_t e0,e1;cudaEventCreate(&e0);cudaEventCreate(&e1);
  int blocks=148*blocks_per_sm; size_t smem=(size_t)n_rec*sizeof(HalfRec);
  cudaFuncSetAttribute(ub<NT,PPT,MINB,VAR>,cudaFuncAttributeMaxDynamicSharedMemorySize,(int)smem);
  ub<NT,PPT,MINB,VAR><<<blocks,NT,smem>>>(n_rec,reps/4,sink);
  cudaEventRecord(e0);
  ub<NT,PPT,MINB,VAR><<<blocks,NT,smem>>>(n_rec,reps,sink);
  cudaEventRecord(e1);cudaEventSynchronize(e1);
  float ms;cudaEventElapsedTime(&ms,e0,e1);
  cudaFuncAttributes fa; cudaFuncGetAttributes(&fa,ub<NT,PPT,MINB,VAR>);
  double evals=(double)reps*n_rec*PPT*blocks*NT;
  printf("NT %d PPT %d CTAs/SM %d var %d regs %d: %.3e evals/s  (%s)\n",NT,PPT,blocks_per_sm,VAR,fa.numRegs,evals/(ms*1e-3),cudaGetErrorString(cudaGetLastError()));
  cudaFree(sink);
}
int main(){
  const int nr=256, reps=400;
  run<128,4,4,0>(4,nr,reps);   // the kernel's configuration
  run<128,4,4,1>(4,nr,reps);
  run<128,4,4,2>(4,nr,reps);
  run<128,4,4,3>(4,nr,reps);
  run<128,4,4,10>(4,nr,reps);  // partial records, per-k vote + branch (the kernel's form); evals/s counts ALL points
  run<128,4,4,11>(4,nr,reps);  // partial records, unconditional evaluation + select
  run<128,4,4,12>(4,nr,reps);  // record-level vote, then unconditional
  run<128,4,4,20>(4,nr,reps);  // FP32 K, FP64 accumulation
  run<128,4,6,20>(6,nr,reps);
  run<128,8,4,20>(4,nr,reps);
  run<128,4,3,0>(3,nr,reps);
  run<128,2,6,0>(6,nr,reps);
  run<128,2,8,0>(8,nr,reps);
  run<256,2,4,0>(4,nr,reps);
  run<128,8,3,0>(3,nr,reps);
  run<128,8,2,0>(2,nr,reps);
  run<64,8,6,0>(6,nr,reps);
  run<128,1,8,0>(8,nr,reps);
  run<256,1,8,0>(8,nr,reps);
  return 0;
}
