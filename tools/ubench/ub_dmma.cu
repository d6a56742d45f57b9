// Micro-benchmark: FP64 tensor-core mma.sync (DMMA) throughput on sm_100a, alone and interleaved
// with DFMA, to decide whether the K3a contraction / the K1 accumulations can use it.
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -cudart shared -o ub_dmma ub_dmma.cu
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void mma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void mma16816(double (&c)[4], const double (&a)[8], const double (&b)[4]) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, "
                 "{%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                   "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}
__device__ __forceinline__ void mma1688(double (&c)[4], const double (&a)[4], const double (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, "
                 "{%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}
// MODE 0: m8n8k4 only; 1: DFMA only; 2: m8n8k4 + NF DFMA per mma; 3: m16n8k16; 4: m16n8k8
template <int ILP, int MODE, int NF>
__global__ void __launch_bounds__(256) ub(int iters, double a0, double b0, double* sink) {
    double c[ILP][4], f[8];
    double a[8], b[4];
#pragma unroll
    for (int i = 0; i < 8; i++) { a[i] = a0 + i * 1e-3 + threadIdx.x * 1e-6; f[i] = i; }
#pragma unroll
    for (int i = 0; i < 4; i++) b[i] = b0 + i * 1e-3;
#pragma unroll
    for (int k = 0; k < ILP; k++) c[k][0] = c[k][1] = c[k][2] = c[k][3] = 0.0;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int k = 0; k < ILP; k++) {
            if (MODE == 0 || MODE == 2) mma884(c[k][0], c[k][1], a[k & 7], b[k & 3]);
            if (MODE == 3) mma16816(c[k], a, b);
            if (MODE == 4) { double a4[4] = {a[0], a[1], a[2], a[3]}; double b2[2] = {b[0], b[1]}; mma1688(c[k], a4, b2); }
            if (MODE == 1 || MODE == 2) {
#pragma unroll
                for (int q = 0; q < NF; q++) f[(k * NF + q) & 7] = fma(f[(k * NF + q) & 7], a[q & 7], b[q & 3]);
            }
        }
    }
    double s = 0;
#pragma unroll
    for (int k = 0; k < ILP; k++) s += c[k][0] + c[k][1] + c[k][2] + c[k][3];
#pragma unroll
    for (int i = 0; i < 8; i++) s += f[i];
    if (s == 1.2345) sink[0] = s;
}
template <int ILP, int MODE, int NF>
void run(const char* name, int bps, int iters) {
    double* sink; cudaMalloc(&sink, 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int blocks = 148 * bps, threads = 256;
    ub<ILP, MODE, NF><<<blocks, threads>>>(iters / 4, 1.0, 0.5, sink);
    cudaEventRecord(e0);
    ub<ILP, MODE, NF><<<blocks, threads>>>(iters, 1.0, 0.5, sink);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double warps = (double)blocks * threads / 32;
    double mma_fma = (MODE == 0 || MODE == 2) ? 256.0 : (MODE == 3 ? 2048.0 : (MODE == 4 ? 1024.0 : 0.0));
    double fma_fma = (MODE == 1 || MODE == 2) ? 32.0 * NF : 0.0;
    double tf_mma = 2 * mma_fma * warps * iters * ILP / (ms * 1e-3) / 1e12;
    double tf_fma = 2 * fma_fma * warps * iters * ILP / (ms * 1e-3) / 1e12;
    printf("%-28s ILP %d bps %d: %.3f ms  DMMA %.2f TF  DFMA %.2f TF  sum %.2f TF (%s)\n", name, ILP, bps, ms,
           tf_mma, tf_fma, tf_mma + tf_fma, cudaGetErrorString(cudaGetLastError()));
    cudaFree(sink);
}
int main() {
    int it = 20000;
    run<1, 0, 0>("m8n8k4", 4, it); run<2, 0, 0>("m8n8k4", 4, it); run<4, 0, 0>("m8n8k4", 4, it); run<8, 0, 0>("m8n8k4", 4, it);
    run<8, 0, 0>("m8n8k4", 8, it); run<8, 0, 0>("m8n8k4", 2, it);
    run<8, 1, 8>("dfma only", 4, it); run<8, 1, 8>("dfma only", 8, it);
    run<8, 2, 1>("m8n8k4 + 1 dfma", 4, it); run<8, 2, 2>("m8n8k4 + 2 dfma", 4, it); run<8, 2, 4>("m8n8k4 + 4 dfma", 4, it);
    run<8, 2, 8>("m8n8k4 + 8 dfma", 4, it);
    run<1, 3, 0>("m16n8k16", 4, it); run<4, 3, 0>("m16n8k16", 4, it); run<8, 3, 0>("m16n8k16", 4, it);
    run<4, 4, 0>("m16n8k8", 4, it); run<8, 4, 0>("m16n8k8", 4, it);
    return 0;
}
