// register-tile study: R rows x 16 chains per thread, x from smem via LDS.128 (pairs of adjacent rows), beta via LDS.128 broadcast
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <cmath>
#include <cuda_runtime.h>
#include <math_constants.h>
#include "../nimbleapt_b200/csrc/engine/apt_dists.cuh"
using namespace apt;
constexpr int P = 50, TB = 16;
// MODE 0 dot only, 1 softplus only, 2 both.  R = rows per thread (even).  acc kept in smem when ACCS.
template <int MODE, int R, int THREADS, int MINB, bool ACCS, int SPB>
__global__ void __launch_bounds__(THREADS, MINB) k(const double* __restrict__ th, const double* __restrict__ xg, double* out, int iters) {
    extern __shared__ __align__(16) double sm[];
    double* thT = sm;                         // [P][TB]
    double* xbuf = sm + P * TB;               // [16 cols][R/2 halves][THREADS*2]
    double* accs = xbuf + 16 * R * THREADS;   // [TB][THREADS]
    const int tid = threadIdx.x;
    for (int i = tid; i < P * TB; i += THREADS) thT[i] = th[i];
    for (int i = tid; i < 16 * R * THREADS; i += THREADS) xbuf[i] = xg[i % 4096];
    double acc[TB];
#pragma unroll
    for (int c = 0; c < TB; c++) { acc[c] = 0.0; if (ACCS) accs[c * THREADS + tid] = 0.0; }
    __syncthreads();
    for (int it = 0; it < iters; it++) {
        double ip[R][TB];
#pragma unroll
        for (int r = 0; r < R; r++)
#pragma unroll
            for (int c = 0; c < TB; c++) ip[r][c] = (MODE == 1) ? (ACCS ? accs[c * THREADS + tid] : acc[c]) * 1e-3 + 0.01 * c + 1e-3 * r : 0.0;
        if (MODE != 1) {
#pragma unroll 10
            for (int j = 0; j < P; j++) {
                double xr[R];
#pragma unroll
                for (int h = 0; h < R / 2; h++) {
                    const double2 x2 = *reinterpret_cast<const double2*>(&xbuf[((((j + it) & 15) * (R / 2) + h) * THREADS + tid) * 2]);
                    xr[2 * h] = x2.x; xr[2 * h + 1] = x2.y;
                }
#pragma unroll
                for (int c = 0; c < TB; c += 2) {
                    const double2 b2 = *reinterpret_cast<const double2*>(&thT[j * TB + c]);
#pragma unroll
                    for (int r = 0; r < R; r++) { ip[r][c] = fma(xr[r], b2.x, ip[r][c]); ip[r][c + 1] = fma(xr[r], b2.y, ip[r][c + 1]); }
                }
            }
        }
        if (MODE != 0) {
#pragma unroll
            for (int r = 0; r < R; r++) {
                if (ACCS) {
                    double t[TB];
#pragma unroll
                    for (int c = 0; c < TB; c++) t[c] = 0.0;
                    bern_logit_acc<TB>((it & 1) ? 1.0 : 0.0, ip[r], t);
#pragma unroll
                    for (int c = 0; c < TB; c++) accs[c * THREADS + tid] += t[c];
                } else bern_logit_acc<TB>((it & 1) ? 1.0 : 0.0, ip[r], acc);
            }
        } else {
#pragma unroll
            for (int r = 0; r < R; r++)
#pragma unroll
                for (int c = 0; c < TB; c++) acc[c] += ip[r][c];
        }
    }
    double s = 0;
#pragma unroll
    for (int c = 0; c < TB; c++) s += ACCS ? accs[c * THREADS + tid] : acc[c];
    out[blockIdx.x * THREADS + tid] = s;
}
template <int MODE, int R, int THREADS, int MINB, bool ACCS>
void run(const char* name, const double* th, const double* xg, double* out) {
    const int grid = 148 * MINB;
    size_t smem = (P * TB + 16 * R * THREADS + TB * THREADS) * 8;
    auto fn = k<MODE, R, THREADS, MINB, ACCS, 8>;
    cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int occ = 0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fn, THREADS, smem);
    cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, fn);
    const double rows_per_iter = (double)grid * THREADS * R;
    const int iters = (int)ceil(1e6 / rows_per_iter);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    fn<<<grid, THREADS, smem>>>(th, xg, out, iters);
    cudaEventRecord(e0);
    for (int i = 0; i < 10; i++) fn<<<grid, THREADS, smem>>>(th, xg, out, iters);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("%-34s R=%d thr=%d occ %d regs %d spill %zu iters %d: %.1f us/pass, %.1f us per 1e6 rows (%s)\n", name, R, THREADS, occ, fa.numRegs, (size_t)fa.localSizeBytes, iters, ms * 100, ms * 100 * 1e6 / (rows_per_iter * iters), cudaGetErrorString(cudaGetLastError()));
}
int main() {
    double *th, *xg, *out;
    cudaMalloc(&th, P * TB * 8); cudaMalloc(&xg, 4096 * 8); cudaMalloc(&out, 1 << 22);
    double h[4096]; for (int i = 0; i < 4096; i++) h[i] = 0.3 * sin(i * 0.37);
    cudaMemcpy(xg, h, sizeof h, cudaMemcpyHostToDevice); cudaMemcpy(th, h, P * TB * 8, cudaMemcpyHostToDevice);
    run<0, 2, 256, 2, false>("dot R2 2cta", th, xg, out);
    run<2, 2, 256, 2, false>("both R2 2cta", th, xg, out);
    run<2, 2, 256, 2, true>("both R2 2cta acc-smem", th, xg, out);
    run<0, 4, 256, 1, false>("dot R4 1cta", th, xg, out);
    run<2, 4, 256, 1, false>("both R4 1cta", th, xg, out);
    run<2, 4, 256, 1, true>("both R4 1cta acc-smem", th, xg, out);
    run<0, 4, 128, 2, false>("dot R4 2cta x128", th, xg, out);
    run<2, 4, 128, 2, false>("both R4 2cta x128", th, xg, out);
    run<2, 4, 128, 2, true>("both R4 2cta x128 acc-smem", th, xg, out);
    run<0, 4, 128, 3, false>("dot R4 3cta x128", th, xg, out);
    run<2, 4, 128, 3, true>("both R4 3cta x128 acc-smem", th, xg, out);
    run<1, 4, 256, 1, false>("softplus R4 1cta", th, xg, out);
    run<1, 2, 256, 2, false>("softplus R2 2cta", th, xg, out);
    run<0, 6, 128, 2, false>("dot R6 2cta x128", th, xg, out);
    return 0;
}
