#include <cstdio>
#include <cuda_runtime.h>
// Is the FP64 tensor-core path (mma.sync.m8n8k4.f64, "DMMA") a second pipe next to DFMA on sm_100a?
// MODE 0: DMMA only, 1: DFMA only, 2: both interleaved (NM DMMAs + NF DFMAs per step, all independent chains).
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
template <int NM, int NF, int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) k(double* out, int iters, double a, double b) {
    double c[NM > 0 ? NM : 1][2], f[NF > 0 ? NF : 1];
#pragma unroll
    for (int i = 0; i < NM; i++) { c[i][0] = threadIdx.x * 1e-3 + i; c[i][1] = i; }
#pragma unroll
    for (int i = 0; i < NF; i++) f[i] = threadIdx.x * 1e-3 + i;
    const double x = a + threadIdx.x * 1e-9, y = b;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
#pragma unroll
            for (int i = 0; i < (NM > NF ? NM : NF); i++) {
                if (i < NM) dmma(c[i][0], c[i][1], x, y);
                if (i < NF) f[i] = fma(f[i], x, y);
            }
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NM; i++) s += c[i][0] + c[i][1];
#pragma unroll
    for (int i = 0; i < NF; i++) s += f[i];
    out[blockIdx.x * THREADS + threadIdx.x] = s;
}
template <int NM, int NF, int THREADS, int MINB>
void run() {
    double* out; cudaMalloc(&out, 1 << 24);
    const int grid = 148 * MINB, iters = 1000;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<NM, NF, THREADS, MINB><<<grid, THREADS>>>(out, iters, 0.999, 1e-3);
    cudaEventRecord(e0);
    k<NM, NF, THREADS, MINB><<<grid, THREADS>>>(out, iters, 0.999, 1e-3);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double warps = (double)grid * THREADS / 32;
    const double fl_m = warps * iters * 8.0 * NM * 512.0, fl_f = (double)grid * THREADS * iters * 8.0 * NF * 2.0;
    printf("NM %2d NF %2d threads %4d ctas/SM %d: %.3f ms  DMMA %.2f TFLOP/s  DFMA %.2f TFLOP/s  sum %.2f\n", NM, NF, THREADS, MINB, ms,
           fl_m / (ms * 1e-3) / 1e12, fl_f / (ms * 1e-3) / 1e12, (fl_m + fl_f) / (ms * 1e-3) / 1e12);
    cudaFree(out);
}
int main() {
    run<8, 0, 256, 2>(); run<4, 0, 256, 2>(); run<2, 0, 256, 2>(); run<8, 0, 512, 2>();
    run<0, 8, 256, 2>(); run<0, 16, 256, 2>();
    run<8, 8, 256, 2>(); run<4, 16, 256, 2>(); run<2, 16, 256, 2>(); run<8, 16, 256, 2>(); run<1, 16, 256, 2>(); run<4, 8, 512, 2>();
    return 0;
}
