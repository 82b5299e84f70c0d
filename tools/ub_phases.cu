// microbenchmark of the grid kernel's compute phases (no HBM streaming): dot (x row times 16 chains) and softplus
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <cmath>
#include <cuda_runtime.h>
#include <math_constants.h>
#include "../nimbleapt_b200/csrc/engine/apt_dists.cuh"
using namespace apt;
constexpr int P = 50, TB = 16, THREADS = 256;
__constant__ double cth[P * TB];
template <int MODE, int R, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) k(const double* __restrict__ th, const double* __restrict__ xg, double* out, int rows_per_thread) {
    extern __shared__ __align__(16) double sm[];
    double* thT = sm;                 // [P][TB]
    double* xbuf = sm + P * TB;       // [10*R][THREADS]
    const int tid = threadIdx.x;
    for (int i = tid; i < P * TB; i += THREADS) thT[i] = th[i];
    for (int i = 0; i < 16 * R; i++) xbuf[i * THREADS + tid] = xg[(i * THREADS + tid) % 4096];
    __syncthreads();
    double acc[TB];
#pragma unroll
    for (int c = 0; c < TB; c++) acc[c] = 0.0;
    for (int it = 0; it < rows_per_thread; it += R) {
        asm volatile("" ::: "memory");
        double ip[R][TB];
#pragma unroll
        for (int r = 0; r < R; r++)
#pragma unroll
            for (int c = 0; c < TB; c++) ip[r][c] = (MODE == 1) ? acc[c] * 1e-3 + 0.01 * c + 1e-3 * r : 0.0;
        if (MODE == 3 || MODE == 4) {
#pragma unroll
            for (int b = 0; b < P / 5; b++) {
                double xr[R][5];
#pragma unroll
                for (int r = 0; r < R; r++)
#pragma unroll
                    for (int e = 0; e < 5; e++) xr[r][e] = xbuf[((((b & 1) * 5 + e + it) & 15) * R + r) * THREADS + tid];
#pragma unroll
                for (int e = 0; e < 5; e++)
#pragma unroll
                    for (int c = 0; c < TB; c++)
#pragma unroll
                        for (int r = 0; r < R; r++) ip[r][c] = fma(xr[r][e], cth[(b * 5 + e) * TB + c], ip[r][c]);
            }
        } else if (MODE != 1) {
#pragma unroll
            for (int b = 0; b < P / 5; b++) {
                double xr[R][5];
#pragma unroll
                for (int r = 0; r < R; r++)
#pragma unroll
                    for (int e = 0; e < 5; e++) xr[r][e] = xbuf[((((b & 1) * 5 + e + it) & 15) * R + r) * THREADS + tid];
#pragma unroll
                for (int e = 0; e < 5; e++)
#pragma unroll
                    for (int c = 0; c < TB; c += 2) {
                        const double2 b2 = *reinterpret_cast<const double2*>(&thT[(b * 5 + e) * TB + c]);
#pragma unroll
                        for (int r = 0; r < R; r++) { ip[r][c] = fma(xr[r][e], b2.x, ip[r][c]); ip[r][c + 1] = fma(xr[r][e], b2.y, ip[r][c + 1]); }
                    }
            }
        }
        if (MODE != 0 && MODE != 3) {
#pragma unroll
            for (int r = 0; r < R; r++) bern_logit_acc<TB>((it & 1) ? 1.0 : 0.0, ip[r], acc);
        } else {
#pragma unroll
            for (int r = 0; r < R; r++)
#pragma unroll
                for (int c = 0; c < TB; c++) acc[c] += ip[r][c];
        }
    }
    double s = 0;
#pragma unroll
    for (int c = 0; c < TB; c++) s += acc[c];
    out[blockIdx.x * THREADS + tid] = s;
}
__global__ void wr(double* c, double v) { for (int i = threadIdx.x; i < P * TB; i += blockDim.x) c[i] = v + i; }
__global__ void rd(double v, int* bad) { int nb = 0;
#pragma unroll 1
  for (int i = 0; i < P * TB; i++) if (cth[i] != v + i) nb++;
  if (nb) atomicAdd(bad, 1); }
template <int MODE, int R, int MINB>
void run(const char* name, int grid, const double* th, const double* xg, double* out) {
    size_t smem = (P * TB + 16 * R * THREADS) * 8;
    cudaFuncSetAttribute(k<MODE, R, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int occ = 0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k<MODE, R, MINB>, THREADS, smem);
    const int rows = 1000000 / (grid * THREADS) + 1;  // same rows per thread as the real kernel at N = 1e6
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE, R, MINB><<<grid, THREADS, smem>>>(th, xg, out, rows + (rows % R));
    cudaEventRecord(e0);
    for (int i = 0; i < 10; i++) k<MODE, R, MINB><<<grid, THREADS, smem>>>(th, xg, out, rows + (rows % R));
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("%-28s grid %d occ %d rows/thread %d: %.1f us per pass-equivalent (%s)\n", name, grid, occ, rows, ms * 100, cudaGetErrorString(cudaGetLastError()));
}
int main() {
    double *th, *xg, *out;
    cudaMalloc(&th, P * TB * 8); cudaMalloc(&xg, 4096 * 8); cudaMalloc(&out, 1 << 22);
    double h[4096]; for (int i = 0; i < 4096; i++) h[i] = 0.3 * sin(i * 0.37);
    cudaMemcpy(xg, h, sizeof h, cudaMemcpyHostToDevice); cudaMemcpy(th, h, P * TB * 8, cudaMemcpyHostToDevice);
    run<0, 1, 2>("dot only R=1 minb2", 296, th, xg, out);
    run<1, 1, 2>("softplus only R=1 minb2", 296, th, xg, out);
    run<2, 1, 2>("both R=1 minb2", 296, th, xg, out);
    run<2, 1, 1>("both R=1 minb1 (1 CTA/SM)", 148, th, xg, out);
    run<2, 1, 3>("both R=1 minb3", 444, th, xg, out);
    run<0, 2, 2>("dot only R=2 minb2", 296, th, xg, out);
    run<2, 2, 2>("both R=2 minb2", 296, th, xg, out);
    run<0, 1, 3>("dot only R=1 minb3", 444, th, xg, out);
    run<1, 1, 3>("softplus only R=1 minb3", 444, th, xg, out);
    run<0, 1, 4>("dot only R=1 minb4", 592, th, xg, out);
    run<1, 1, 4>("softplus only R=1 minb4", 592, th, xg, out);
    cudaMemcpyToSymbol(cth, h, P * TB * 8);
    run<3, 1, 2>("dot only CONST R=1 minb2", 296, th, xg, out);
    run<4, 1, 2>("both CONST R=1 minb2", 296, th, xg, out);
    run<3, 1, 3>("dot only CONST R=1 minb3", 444, th, xg, out);
    run<4, 1, 3>("both CONST R=1 minb3", 444, th, xg, out);
    run<3, 2, 2>("dot only CONST R=2 minb2", 296, th, xg, out);
    run<4, 2, 2>("both CONST R=2 minb2", 296, th, xg, out);
    // visibility of device-written constant memory to the next launch
    {
        double* cptr; cudaGetSymbolAddress((void**)&cptr, cth);
        int* bad; cudaMalloc(&bad, 4); cudaMemset(bad, 0, 4);
        for (int i = 0; i < 20000; i++) { wr<<<1, 256>>>(cptr, (double)i); rd<<<148, 256>>>((double)i, bad); }
        int hb = -1; cudaMemcpy(&hb, bad, 4, cudaMemcpyDeviceToHost);
        printf("device-written constant visibility: %d stale reads in 20000 writer->reader launches (%s)\n", hb, cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
