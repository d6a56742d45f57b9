// Micro-benchmark: does a DFMA (2 pipe cycles per warp instruction per SM sub-partition) leave its
// second cycle free for another instruction?  8 independent DFMA chains per thread plus K extra
// integer (IMAD/LOP) or shared-memory (LDS) instructions per 8 DFMA.  If the extras were free in
// the DFMA shadow the time would not change up to K = 8.
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -cudart shared -o ub_issue ub_issue.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int K, int KIND>
__global__ void __launch_bounds__(256) ub(int iters, double m, double c, int im, double* sink) {
    __shared__ double sh[256];
    sh[threadIdx.x] = threadIdx.x;
    __syncthreads();
    double a[8];
#pragma unroll
    for (int k = 0; k < 8; k++) a[k] = threadIdx.x + k;
    int x[8];
#pragma unroll
    for (int k = 0; k < 8; k++) x[k] = threadIdx.x * 3 + k;
    double ld = 0.0;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int k = 0; k < 8; k++) {
            a[k] = fma(a[k], m, c);
            if (k < K) {
                if (KIND == 0) x[k] = x[k] * im + i;                       // IMAD
                else if (KIND == 1) x[k] = (x[k] ^ i) & im;                // LOP3
                else { ld += sh[(x[k] + i) & 255]; }                       // LDS + address + DADD
            }
        }
    }
    double s = ld;
    int t = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) { s += a[k]; t += x[k]; }
    if (s == 1.2345 || t == 12345) sink[0] = s + t;
}
template <int K, int KIND>
void run(int iters) {
    double* sink;
    cudaMalloc(&sink, 8);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int blocks = 148 * 8;
    ub<K, KIND><<<blocks, 256>>>(iters / 4, 0.999999, 1e-9, 3, sink);
    cudaEventRecord(e0);
    ub<K, KIND><<<blocks, 256>>>(iters, 0.999999, 1e-9, 3, sink);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double dfma = 8.0 * iters * blocks * 256;
    printf("kind %d (0 IMAD, 1 LOP3, 2 LDS+DADD)  extras per 8 DFMA = %d : %.3f ms, %.2f TFLOP/s FP64\n", KIND, K,
           ms, 2.0 * dfma / (ms * 1e-3) / 1e12);
    cudaFree(sink);
}
int main() {
    const int it = 20000;
    run<0, 0>(it); run<2, 0>(it); run<4, 0>(it); run<8, 0>(it);
    run<2, 1>(it); run<4, 1>(it); run<8, 1>(it);
    run<2, 2>(it); run<4, 2>(it); run<8, 2>(it);
    return 0;
}
