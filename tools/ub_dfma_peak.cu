#include <cstdio>
#include <cuda_runtime.h>
// pure DFMA throughput: NACC independent accumulators per thread, operands in registers
template <int NACC, int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) k(double* out, int iters, double a, double b) {
    double acc[NACC];
#pragma unroll
    for (int i = 0; i < NACC; i++) acc[i] = threadIdx.x * 1e-3 + i;
    double x = a + threadIdx.x * 1e-9, y = b;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 16; u++) {
#pragma unroll
            for (int i = 0; i < NACC; i++) acc[i] = fma(acc[i], x, y);
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; i++) s += acc[i];
    out[blockIdx.x * THREADS + threadIdx.x] = s;
}
template <int NACC, int THREADS, int MINB>
void run() {
    double* out; cudaMalloc(&out, 1 << 24);
    const int grid = 148 * MINB, iters = 2000;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<NACC, THREADS, MINB><<<grid, THREADS>>>(out, iters, 0.999, 1e-3);
    cudaEventRecord(e0);
    k<NACC, THREADS, MINB><<<grid, THREADS>>>(out, iters, 0.999, 1e-3);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double fma = (double)grid * THREADS * iters * 16.0 * NACC;
    printf("NACC %2d threads %4d ctas/SM %d: %.2f TFLOP/s, %.1f DFMA/clk/SM @1.9GHz\n", NACC, THREADS, MINB, 2 * fma / (ms * 1e-3) / 1e12, fma / (ms * 1e-3) / 148 / 1.9e9);
    cudaFree(out);
}
int main() {
    run<1, 256, 2>(); run<2, 256, 2>(); run<4, 256, 2>(); run<8, 256, 2>(); run<16, 256, 2>(); run<32, 256, 2>();
    run<8, 128, 1>(); run<16, 128, 1>(); run<32, 128, 1>(); run<8, 1024, 2>(); run<4, 1024, 2>();
    return 0;
}
