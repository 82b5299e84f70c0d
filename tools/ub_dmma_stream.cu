// Streaming-pass study for config 3 (N = 1e6 rows, p = 50, 16 rungs): the X beta product of a 32-row warp tile on the FP64
// tensor-core path (mma.sync.m8n8k4.f64) instead of 800 DFMA + 200 LDS.128 per tile.  DMMA and DFMA share one pipe on
// sm_100a (tools/ub_dmma.cu), so the FP64 floor does not move; what DMMA saves is issue slots and operand loads.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o ub_dmma_stream tools/ub_dmma_stream.cu && ./ub_dmma_stream
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>
#include <math_constants.h>
#include "../nimbleapt_b200/csrc/engine/apt_dists.cuh"
using namespace apt;
constexpr int P = 50, TB = 16, KB = 4, NBR = (P + KB - 1) / KB;
#ifndef RB
#define RB 4  // row blocks of 8 rows per warp tile
#endif
constexpr int TRW = 8 * RB, XS = TRW + 4;  // XS: padded column stride of a ring block (doubles): LDS.64 of a fragment hits 16 distinct bank pairs

__device__ __forceinline__ void cpa16(void* smem, const void* gmem) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cpa_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cpa_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void mbar_init(uint64_t* b, unsigned n) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(b)), "r"(n) : "memory"); }
__device__ __forceinline__ void mbar_expect(uint64_t* b, unsigned bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, uint64_t* b) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src), "r"(bytes), "r"((unsigned)__cvta_generic_to_shared(b)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, unsigned parity) {
    asm volatile("{\n.reg .pred p;\nWAIT_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}" ::"r"((unsigned)__cvta_generic_to_shared(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// BREG: B fragments (26 doubles) live in registers for the whole kernel; else re-read from shared memory per block.
// DEPTH: blocks in flight (ring slots = DEPTH + 1).  MODE 0: dot only, 2: dot + Bernoulli-logit terms.
template <int THREADS, int MINB, bool BREG, int DEPTH, int MODE, int UNR = 1, bool BULK = false>
__global__ void __launch_bounds__(THREADS, MINB) pass_kernel(const double* __restrict__ X, const double* __restrict__ y, const double* __restrict__ th, long n, double* partial) {
    constexpr int NW = THREADS / 32, SLOTS = DEPTH + 1, BLK = KB * XS;
    extern __shared__ __align__(16) double sm[];
    double* thT = sm;                       // [NBR*KB][TB] param-major, zero rows beyond P
    double* red = thT + NBR * KB * TB;      // [NW][TB]
    double* xbuf = red + NW * TB;           // [NW][SLOTS][KB][XS]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < NBR * KB * TB; i += THREADS) thT[i] = (i < P * TB) ? th[i] : 0.0;
    double* xw = xbuf + warp * SLOTS * BLK;
    const long ntile = n / TRW, wstep = (long)gridDim.x * NW;
    long wt = (long)warp * gridDim.x + blockIdx.x;
    // B fragment of block b, n-tile t: B[k = lane % 4][col = lane / 4]; column col of n-tile t is position (col/2)*4 + t*2 + col%2
    const int kq = lane & 3, rq = lane >> 2;
    const int bpos0 = (rq >> 1) * 4 + (rq & 1);
    uint64_t* bars = reinterpret_cast<uint64_t*>(xbuf + NW * SLOTS * BLK) + warp * SLOTS;
    if (BULK) {
        if (lane == 0) for (int i = 0; i < SLOTS; i++) mbar_init(&bars[i], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        __syncwarp();
    }
    unsigned phases = 0;  // parity of each slot's barrier
    auto issue = [&](long tile, int b, int slot) {
        if (BULK) {
            if (lane == 0) {
                const int nc = (b * KB + KB <= P) ? KB : P - b * KB;
                mbar_expect(&bars[slot], (unsigned)nc * (unsigned)TRW * 8u);
                for (int col = 0; col < nc; col++) bulk_g2s(&xw[slot * BLK + col * XS], X + tile * TRW + (long)(b * KB + col) * n, (unsigned)TRW * 8u, &bars[slot]);
            }
            return;
        }
        // block b = columns 4b .. 4b+3 of rows tile*32 .. +31: 4 x 16 chunks of 16 bytes, 2 per lane
        const double* src = X + tile * TRW;
        constexpr int CPC = TRW / 2;  // 16-byte chunks per column
#pragma unroll
        for (int j = 0; j < (KB * CPC + 31) / 32; j++) {
            const int c = lane + 32 * j, col = c / CPC, q = c % CPC;
            if (c < KB * CPC && b * KB + col < P) cpa16(&xw[slot * BLK + col * XS + 2 * q], src + (long)(b * KB + col) * n + 2 * q);
        }
    };
    if (wt < ntile) {
#pragma unroll
        for (int b = 0; b < DEPTH; b++) { issue(wt, b, b); if (!BULK) cpa_commit(); }
    }
    __syncthreads();
    double breg[BREG ? NBR : 1][2];
    if (BREG) {
#pragma unroll
        for (int b = 0; b < NBR; b++) { breg[b][0] = thT[(b * KB + kq) * TB + bpos0]; breg[b][1] = thT[(b * KB + kq) * TB + bpos0 + 2]; }
    }
    double acc[4];
#pragma unroll
    for (int c = 0; c < 4; c++) acc[c] = 0.0;
    int ring = 0;
    for (; wt < ntile; wt += wstep) {
        double c[RB][2][2];  // [row block][n-tile][2]
#pragma unroll
        for (int i = 0; i < RB; i++) { c[i][0][0] = c[i][0][1] = c[i][1][0] = c[i][1][1] = 0.0; }
        const long nxt = wt + wstep;
#pragma unroll (BREG ? NBR : UNR)
        for (int b = 0; b < NBR; b++) {
            if (BULK) { mbar_wait(&bars[ring], (phases >> ring) & 1u); phases ^= 1u << ring; }
            else cpa_wait<DEPTH - 1>();
            __syncwarp();
            {
                const int nb = b + DEPTH;
                int slot = ring + DEPTH; if (slot >= SLOTS) slot -= SLOTS;
                if (nb < NBR) issue(wt, nb, slot);
                else if (nxt < ntile) issue(nxt, nb - NBR, slot);
                if (!BULK) cpa_commit();
            }
            double b0, b1;
            if (BREG) { b0 = breg[b][0]; b1 = breg[b][1]; }
            else { b0 = thT[(b * KB + kq) * TB + bpos0]; b1 = thT[(b * KB + kq) * TB + bpos0 + 2]; }
            const double* xs = xw + ring * BLK + kq * XS + rq;
            const bool live = b * KB + kq < P;
#pragma unroll
            for (int i = 0; i < RB; i++) {
                double a = xs[8 * i];
                if (b == NBR - 1) a = live ? a : 0.0;
                dmma(c[i][0][0], c[i][0][1], a, b0);
                dmma(c[i][1][0], c[i][1][1], a, b1);
            }
            ring = (ring + 1 == SLOTS) ? 0 : ring + 1;
        }
        // thread tile: rows rq + 8 i, chains (positions) 4 kq + {0, 1, 2, 3} = n-tile 0 cols 2kq, 2kq+1, n-tile 1 cols 2kq, 2kq+1
        if (MODE == 2) {
            double yy[RB], eta[RB][4];
#pragma unroll
            for (int i = 0; i < RB; i++) {
                yy[i] = y[wt * TRW + rq + 8 * i];
                eta[i][0] = c[i][0][0]; eta[i][1] = c[i][0][1]; eta[i][2] = c[i][1][0]; eta[i][3] = c[i][1][1];
            }
            bern_logit_tile<RB, 4>(yy, eta, acc);
        } else {
#pragma unroll
            for (int i = 0; i < RB; i++) { acc[0] += c[i][0][0]; acc[1] += c[i][0][1]; acc[2] += c[i][1][0]; acc[3] += c[i][1][1]; }
        }
    }
    if (!BULK) cpa_wait<0>();
#pragma unroll
    for (int cc = 0; cc < 4; cc++) {
        double v = acc[cc];
#pragma unroll
        for (int o = 16; o >= 4; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane < 4) red[warp * TB + kq * 4 + cc] = v;
    }
    __syncthreads();
    if (tid < TB) {
        double s = 0.0;
        for (int w = 0; w < NW; w++) s += red[w * TB + tid];
        partial[blockIdx.x * TB + tid] = s;
    }
}

// plain reference: one thread per row, sequential fma chain, library-free terms through the same bern_logit_tile
__global__ void ref_kernel(const double* __restrict__ X, const double* __restrict__ y, const double* __restrict__ th, long n, double* out, int mode) {
    const long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    for (int c0 = 0; c0 < TB; c0 += 4) {
        double eta[1][4] = {{0, 0, 0, 0}};
        for (int j = 0; j < P; j++) {
            const double x = X[(long)j * n + i];
            for (int c = 0; c < 4; c++) eta[0][c] = fma(x, th[j * TB + c0 + c], eta[0][c]);
        }
        double acc[4] = {0, 0, 0, 0}, yy[1] = {y[i]};
        if (mode == 2) bern_logit_tile<1, 4>(yy, eta, acc);
        else for (int c = 0; c < 4; c++) acc[c] = eta[0][c];
        for (int c = 0; c < 4; c++) atomicAdd(&out[c0 + c], acc[c]);
    }
}

template <int THREADS, int MINB, bool BREG, int DEPTH, int MODE, int UNR = 1, bool BULK = false>
void run(const char* name, const double* X, const double* y, const double* th, long n, double* partial, const double* ref) {
    const int grid = 148 * MINB;
    constexpr int NW = THREADS / 32;
    size_t smem = (size_t)(NBR * KB * TB + NW * TB + NW * (DEPTH + 1) * KB * XS + NW * (DEPTH + 1)) * 8;
    auto fn = pass_kernel<THREADS, MINB, BREG, DEPTH, MODE, UNR, BULK>;
    cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int occ = 0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fn, THREADS, smem);
    cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, fn);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int i = 0; i < 3; i++) fn<<<grid, THREADS, smem>>>(X, y, th, n, partial);
    cudaEventRecord(e0);
    const int reps = 20;
    for (int i = 0; i < reps; i++) fn<<<grid, THREADS, smem>>>(X, y, th, n, partial);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    std::vector<double> h(grid * TB);
    cudaMemcpy(h.data(), partial, h.size() * 8, cudaMemcpyDeviceToHost);
    double worst = 0;
    for (int c = 0; c < TB; c++) { double s = 0; for (int b = 0; b < grid; b++) s += h[b * TB + c]; worst = fmax(worst, fabs(s - ref[c]) / fabs(ref[c])); }
    printf("%-28s thr=%d minb=%d occ %d regs %d spill %zu smem %zu: %.1f us/pass (%.3f of 6551.7 GB/s)  max rel err vs ref %.2e (%s)\n", name, THREADS, MINB, occ, fa.numRegs,
           (size_t)fa.localSizeBytes, smem, ms * 1000 / reps, 8.0 * n * (P + 1) / (ms / reps * 1e-3) / 1e9 / 6551.7, worst, cudaGetErrorString(cudaGetLastError()));
}

int main() {
    const long n = 1000000;
    double *X, *y, *th, *partial, *refd;
    cudaMalloc(&X, n * P * 8); cudaMalloc(&y, n * 8); cudaMalloc(&th, P * TB * 8); cudaMalloc(&partial, 148 * 4 * TB * 8); cudaMalloc(&refd, 2 * TB * 8);
    std::vector<double> hx(n * P), hy(n), ht(P * TB);
    uint64_t s = 88172645463325252ull;
    auto rnd = [&]() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return (double)(s >> 11) / 9007199254740992.0; };
    for (auto& v : hx) v = 2.0 * rnd() - 1.0;
    for (auto& v : hy) v = rnd() < 0.5 ? 0.0 : 1.0;
    for (int j = 0; j < P; j++) for (int c = 0; c < TB; c++) ht[j * TB + c] = (rnd() - 0.5) * (0.2 + 0.1 * c);  // hotter rungs: larger |eta|
    cudaMemcpy(X, hx.data(), n * P * 8, cudaMemcpyHostToDevice); cudaMemcpy(y, hy.data(), n * 8, cudaMemcpyHostToDevice); cudaMemcpy(th, ht.data(), P * TB * 8, cudaMemcpyHostToDevice);
    double ref[2][TB];
    for (int mode = 0; mode <= 2; mode += 2) {
        cudaMemset(refd, 0, 2 * TB * 8);
        ref_kernel<<<(n + 255) / 256, 256>>>(X, y, th, n, refd, mode);
        cudaMemcpy(ref[mode / 2], refd, TB * 8, cudaMemcpyDeviceToHost);
    }
    // the reference sums positions as chains directly; pass_kernel's positions are the same columns of th
#if RB == 4
    run<256, 2, true, 3, 0>("dot breg d3", X, y, th, n, partial, ref[0]);
    run<256, 2, false, 3, 0>("dot blds d3", X, y, th, n, partial, ref[0]);
    run<256, 2, true, 3, 2>("both breg d3", X, y, th, n, partial, ref[1]);
    run<256, 2, false, 3, 2>("both blds d3", X, y, th, n, partial, ref[1]);
    run<256, 2, false, 4, 2>("both blds d4", X, y, th, n, partial, ref[1]);
    run<256, 2, false, 6, 2>("both blds d6", X, y, th, n, partial, ref[1]);
    run<256, 3, false, 3, 2>("both blds d3 3cta", X, y, th, n, partial, ref[1]);
    run<256, 3, false, 4, 2>("both blds d4 3cta", X, y, th, n, partial, ref[1]);
    run<128, 4, false, 3, 2>("both blds d3 4x128", X, y, th, n, partial, ref[1]);
    run<128, 6, false, 3, 2>("both blds d3 6x128", X, y, th, n, partial, ref[1]);
    run<512, 1, false, 3, 2>("both blds d3 1x512", X, y, th, n, partial, ref[1]);
    run<128, 4, true, 3, 2>("both breg d3 4x128", X, y, th, n, partial, ref[1]);
    run<256, 2, false, 3, 2, 13>("both blds d3 unr13", X, y, th, n, partial, ref[1]);
    run<256, 2, false, 3, 2, 2>("both blds d3 unr2", X, y, th, n, partial, ref[1]);
    run<256, 2, false, 6, 2, 13>("both blds d6 unr13", X, y, th, n, partial, ref[1]);
    run<256, 2, false, 3, 2, 13, true>("both blds d3 unr13 BULK", X, y, th, n, partial, ref[1]);
    run<256, 2, false, 6, 2, 13, true>("both blds d6 unr13 BULK", X, y, th, n, partial, ref[1]);
    run<256, 2, false, 6, 0, 13, true>("dot blds d6 unr13 BULK", X, y, th, n, partial, ref[0]);
    run<256, 2, false, 6, 0, 13>("dot blds d6 unr13", X, y, th, n, partial, ref[0]);
#elif RB == 8
    run<256, 1, false, 6, 2, 13>("both d6 unr13 1x256", X, y, th, n, partial, ref[1]);
    run<384, 1, false, 6, 2, 13>("both d6 unr13 1x384", X, y, th, n, partial, ref[1]);
    run<512, 1, false, 4, 2, 13>("both d4 unr13 1x512", X, y, th, n, partial, ref[1]);
    run<256, 2, false, 3, 2, 13>("both d3 unr13 2x256", X, y, th, n, partial, ref[1]);
#else
    run<256, 2, false, 6, 2, 13>("both d6 unr13 2x256", X, y, th, n, partial, ref[1]);
    run<256, 3, false, 6, 2, 13>("both d6 unr13 3x256", X, y, th, n, partial, ref[1]);
    run<256, 3, false, 8, 2, 13>("both d8 unr13 3x256", X, y, th, n, partial, ref[1]);
    run<256, 4, false, 6, 2, 13>("both d6 unr13 4x256", X, y, th, n, partial, ref[1]);
    run<128, 6, false, 6, 2, 13>("both d6 unr13 6x128", X, y, th, n, partial, ref[1]);
    run<128, 8, false, 6, 2, 13>("both d6 unr13 8x128", X, y, th, n, partial, ref[1]);
#endif
    return 0;
}
