// apt_host.cuh — host side of a lowered-model module: device state, the run prologue of buildAPT$run
// (/root/reference/nimbleAPT/R/APT_functions.R:350-393), kernel launches on a private stream, CUDA-event
// timing, getters.  Instantiated once per generated model G and exported through the plain-C module ABI
// (apt_module_abi.h) by APT_DEFINE_MODULE(G).  No torch, no CPU fallback: every failure is an error string.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include "apt_engine.cuh"
#include "apt_grid.cuh"
#include "apt_module_abi.h"
#include "apt_waic.cuh"

namespace apt {

struct HostSamplerInfo {
    int kind, d, unit0, count, blkoff, empoff, adapt_interval;
    double scale0;
    const double* propcov0;  // d*d column-major or nullptr (identity)
};

struct ApiError {
    std::string msg;
};
#define APT_CUDA(x)                                                                                   \
    do {                                                                                              \
        cudaError_t e_ = (x);                                                                         \
        if (e_ != cudaSuccess) throw ApiError{std::string(#x) + ": " + cudaGetErrorString(e_)};      \
    } while (0)

inline int host_chol_upper(const double* A, double* U, int n) {
    for (int i = 0; i < n * n; i++) U[i] = 0.0;
    for (int j = 0; j < n; j++) {
        double s = A[j + j * n];
        for (int k = 0; k < j; k++) s -= U[k + j * n] * U[k + j * n];
        if (!(s > 0)) return -1;
        double d = std::sqrt(s);
        U[j + j * n] = d;
        for (int i = j + 1; i < n; i++) {
            double t = A[j + i * n];
            for (int k = 0; k < j; k++) t -= U[k + j * n] * U[k + i * n];
            U[j + i * n] = t / d;
        }
    }
    return 0;
}

template <class T>
struct DevBuf {
    T* p = nullptr;
    size_t n = 0;
    void alloc(size_t count) {
        release();
        n = count;
        if (count) APT_CUDA(cudaMalloc((void**)&p, count * sizeof(T)));
    }
    void zero(cudaStream_t s) {
        if (n) APT_CUDA(cudaMemsetAsync(p, 0, n * sizeof(T), s));
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        n = 0;
    }
    ~DevBuf() { release(); }
};

// pure FP64 FMA throughput (the flop-roofline denominator): 16 independent accumulators per thread
__global__ void __launch_bounds__(256, 2) fp64_peak_kernel(double* out, int iters, double a, double b) {
    double acc[16];
#pragma unroll
    for (int i = 0; i < 16; i++) acc[i] = threadIdx.x * 1e-3 + i;
    const double x = a + threadIdx.x * 1e-9;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 16; u++) {
#pragma unroll
            for (int i = 0; i < 16; i++) acc[i] = fma(acc[i], x, b);
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// [rows][cols] -> [cols][rows]: the column-major getters (R matrices) transpose on the device, so that the copy lands in the
// caller's matrix as it is
__global__ void transpose_kernel(const double* __restrict__ src, long rows, int cols, double* __restrict__ dst) {
    __shared__ double t[32][33];
    const long r0 = (long)blockIdx.x * 32;
    const int c0 = blockIdx.y * 32;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const long r = r0 + i;
        const int c = c0 + threadIdx.x;
        if (r < rows && c < cols) t[i][threadIdx.x] = src[r * cols + c];
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int c = c0 + i;
        const long r = r0 + threadIdx.x;
        if (r < rows && c < cols) dst[(size_t)c * rows + r] = t[threadIdx.x][i];
    }
}

template <class G>
__global__ void derived_kernel(Data D, int k, long n, double* out) {
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) G::der_fill(D, k, i, out);
}

template <class G>
struct Engine {
    using SH = Shape<G>;
    static constexpr int K = G::K, P = G::P;
    int L = 0, device = 0;
    uint32_t ladder0 = 0, seed = 0;
    double ULT = 1e6;
    int monitor_tmax = 1, record = 0, thin_conf = 1, thin2_conf = 1;
    cudaStream_t stream = nullptr;
    // asynchronous copy-back of the rows a run writes (aptmod_run_args.samples_out): second stream, issued per launch chunk
    cudaStream_t copy_stream = nullptr;
    double* so_ptr = nullptr;
    long long so_nr = 0, so_off = 0;
    int so_thin = 1;
    // device-side hand-off of the same rows (aptmod_run_args.chunk_cb: the NCCL gather lives in the front library)
    int (*chunk_cb)(void*, const double*, int64_t, int64_t, int64_t, void*) = nullptr;
    void* chunk_arg = nullptr;
    DevBuf<double> pack;
    // trace rows [r0, r1) of all L ladders between two [L][pitch rows][cols] layouts; cudaMemcpy2D rejects pitches of 2 GiB
    // and more (cudaDevAttrMaxPitch), so very long traces go ladder by ladder
    static void copy_rows_2d(double* dst, size_t dpitch_d, const double* src, size_t spitch_d, size_t width_d, int nl,
                             cudaMemcpyKind kind, cudaStream_t s) {
        if (width_d == 0 || nl <= 0) return;
        const size_t lim = (size_t)1 << 31;
        if (nl > 1 && dpitch_d * 8 < lim && spitch_d * 8 < lim)
            APT_CUDA(cudaMemcpy2DAsync(dst, dpitch_d * 8, src, spitch_d * 8, width_d * 8, (size_t)nl, kind, s));
        else
            for (int l = 0; l < nl; l++) APT_CUDA(cudaMemcpyAsync(dst + (size_t)l * dpitch_d, src + (size_t)l * spitch_d, width_d * 8, kind, s));
    }
    void copy_rows_async(long long it0, long long it1) {
        if (!so_ptr && !chunk_cb) return;
        const long long r0 = it0 / so_thin, r1 = it1 / so_thin;  // rows of this run written by iterations [it0, it1)
        if (r1 <= r0) return;
        if (chunk_cb) {
            double* dst = pack.p + (size_t)L * r0 * G::NMON;
            copy_rows_2d(dst, (size_t)(r1 - r0) * G::NMON, samples.p + (size_t)(so_off + r0) * G::NMON, (size_t)cap * G::NMON,
                         (size_t)(r1 - r0) * G::NMON, L, cudaMemcpyDeviceToDevice, stream);
            if (chunk_cb(chunk_arg, dst, (int64_t)L * (r1 - r0) * G::NMON, r0, r1, (void*)stream))
                throw ApiError{"the device-side hand-off of the thinned output failed (see the caller's error)"};
            return;
        }
        copy_rows_2d(so_ptr + (size_t)r0 * G::NMON, (size_t)so_nr * G::NMON, samples.p + (size_t)(so_off + r0) * G::NMON,
                     (size_t)cap * G::NMON, (size_t)(r1 - r0) * G::NMON, L, cudaMemcpyDeviceToHost, copy_stream);
    }
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    cudaEvent_t evp[4] = {nullptr, nullptr, nullptr, nullptr};  // start / end of the two pieces in flight
    std::vector<DevBuf<double>> arrays;
    std::vector<DevBuf<int>> iarrays;
    Data D{};
    DevBuf<double> theta, lpd, temps, lp, scale, blk, empir, famdelta, wpart, model_theta, scale0, blk0;
    bool blk_initialised = false;  // sampler$reset() leaves some entries alone (Model::blk_keep) once they exist
    DevBuf<int> slot, cnt, accswap;
    DevBuf<double> g_partial, g_pend, g_scratch, g_stage;  // grid engine
    DevBuf<unsigned int> g_ticket;
    DevBuf<long long> g_dbg, g_delta_tag;
    DevBuf<unsigned long long> g_flags;  // hand-over between pass kernels (apt_grid.cuh)
    unsigned long long launch_seq = 0;
    DevBuf<double> g_delta;
    int num_sms = 148, grid_occ = 1;
    DevBuf<double> samples, samples2, samplesT, samples2T, logprobs, temptraj;
    DevBuf<double> tabZ, tabU, tabS;
    DevBuf<unsigned char> rec;
    DevBuf<int> recswap;
    long long cap = 0, cap2 = 0, ttcap = 0, rows = 0, rows2 = 0, ttrows = 0, rec_iters = 0;
    long long totalIters = 1;
    int last_thin = 1;  // thinToUseVec[1] of the last run (APT:351,605)
    uint64_t rngIter = 0;
    bool resetAnyway = true;
    aptmod_timing timing{};
    long long launches_total = 0;

    void create(const aptmod_init* in) {
        L = in->n_ladders;
        if (L < 1) throw ApiError{"n_ladders must be >= 1"};
        if (in->n_temps != K) throw ApiError{"module was lowered for a different number of temperatures"};
        device = in->device;
        int ndev = 0;
        cudaError_t e = cudaGetDeviceCount(&ndev);
        if (e != cudaSuccess || ndev == 0)
            throw ApiError{std::string("no CUDA device available (aptb200 has no CPU fallback): ") + cudaGetErrorString(e)};
        APT_CUDA(cudaSetDevice(device));
        cudaDeviceProp prop;
        APT_CUDA(cudaGetDeviceProperties(&prop, device));
        if (prop.major != 10)
            throw ApiError{"aptb200 modules are compiled for sm_100a (B200) only; found compute capability " +
                           std::to_string(prop.major) + "." + std::to_string(prop.minor)};
        APT_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
        APT_CUDA(cudaEventCreate(&ev0));
        APT_CUDA(cudaEventCreate(&ev1));
        ladder0 = in->ladder0; seed = in->seed; ULT = in->ULT; monitor_tmax = in->monitor_tmax; record = in->record;
        thin_conf = in->thin_conf > 0 ? in->thin_conf : 1;
        thin2_conf = in->thin2_conf > 0 ? in->thin2_conf : 1;
        if (in->n_arrays != G::NARR || in->n_iarrays != G::NIARR) throw ApiError{"array count mismatch between lowering and module"};
        arrays.resize(G::NARR);
        for (int i = 0; i < G::NARR; i++) {
            arrays[i].alloc((size_t)std::max<int64_t>(in->array_len[i], 1));
            if (in->array_len[i])
                APT_CUDA(cudaMemcpyAsync(arrays[i].p, in->arrays[i], in->array_len[i] * sizeof(double), cudaMemcpyHostToDevice, stream));
            D.a[i] = arrays[i].p;
        }
        iarrays.resize(G::NIARR);
        for (int i = 0; i < G::NIARR; i++) {
            iarrays[i].alloc((size_t)std::max<int64_t>(in->iarray_len[i], 1));
            if (in->iarray_len[i])
                APT_CUDA(cudaMemcpyAsync(iarrays[i].p, in->iarrays[i], in->iarray_len[i] * sizeof(int), cudaMemcpyHostToDevice, stream));
            D.ia[i] = iarrays[i].p;
        }
        refresh_derived();
        const size_t nch = (size_t)L * K;
        theta.alloc(nch * P); lpd.alloc(nch * G::ND); temps.alloc(nch); lp.alloc(nch); slot.alloc(nch);
        scale.alloc(nch * G::NU); cnt.alloc(nch * G::NU * 3);
        blk.alloc(std::max<size_t>(nch * G::BLKSZ, 1)); empir.alloc(std::max<size_t>(nch * G::EMPSZ, 1));
        accswap.alloc(nch * K); model_theta.alloc((size_t)L * P);
        famdelta.alloc(G::HAS_FAMILY ? nch * G::NU * APT_FAM_MAXG : 1); famdelta.zero(stream);
        wpart.alloc((size_t)L * 8 * APT_FAM_MAXG); wpart.zero(stream);
        empir.zero(stream); lp.zero(stream);
        // Temps <- sort(Temps)  (APT:266)
        std::vector<double> t(in->temps, in->temps + K);
        std::sort(t.begin(), t.end());
        std::vector<double> tl(nch);
        for (size_t i = 0; i < nch; i++) tl[i] = t[i % K];
        APT_CUDA(cudaMemcpyAsync(temps.p, tl.data(), nch * sizeof(double), cudaMemcpyHostToDevice, stream));
        std::vector<double> mt((size_t)L * P);
        for (size_t i = 0; i < mt.size(); i++) mt[i] = in->theta0[i % P];
        APT_CUDA(cudaMemcpyAsync(model_theta.p, mt.data(), mt.size() * sizeof(double), cudaMemcpyHostToDevice, stream));
        // sampler reset images (scaleOriginal, propCovOriginal, chol(propCov), scale*chol: APT:675,796-806)
        std::vector<double> s0(G::NU, 1.0), b0(std::max(G::BLKSZ, 1), 0.0);
        for (int s = 0; s < G::NSAMP; s++) {
            const HostSamplerInfo& hi = G::sampler_info(s);
            for (int m = 0; m < hi.count; m++) s0[hi.unit0 + m] = hi.scale0;
            if (hi.kind == 1) {
                const int d = hi.d;
                double* pc = &b0[hi.blkoff];
                for (int i = 0; i < d * d; i++) pc[i] = hi.propcov0 ? hi.propcov0[i] : ((i % (d + 1)) == 0 ? 1.0 : 0.0);
                if (host_chol_upper(pc, pc + d * d, d)) throw ApiError{"propCov is not positive definite"};
                for (int i = 0; i < d * d; i++) pc[2 * d * d + i] = pc[d * d + i] * hi.scale0;
            }
            if (hi.kind == 3) {  // APT:1029-1040: u not drawn yet; five zero matrices, ENSwapMatrix, ENSwapDeltaMatrix, RescaleThreshold
                const int dd = hi.d * hi.d;
                double* b = &b0[hi.blkoff];
                b[0] = -1.0;
                for (int i = 0; i < dd; i++) { b[1 + 5 * dd + i] = 1.0; b[1 + 6 * dd + i] = 1.0; b[1 + 7 * dd + i] = 0.2; }
            }
        }
        scale0.alloc(G::NU); blk0.alloc(b0.size());
        APT_CUDA(cudaMemcpyAsync(scale0.p, s0.data(), s0.size() * sizeof(double), cudaMemcpyHostToDevice, stream));
        APT_CUDA(cudaMemcpyAsync(blk0.p, b0.data(), b0.size() * sizeof(double), cudaMemcpyHostToDevice, stream));
        num_sms = prop.multiProcessorCount;
        if constexpr (G::ENGINE == 2) {
            g_partial.alloc((size_t)num_sms * 8 * 2 * G::TB);
            g_ticket.alloc(1); g_ticket.zero(stream);
            g_dbg.alloc(80); g_dbg.zero(stream);
            g_flags.alloc(2); g_flags.zero(stream);
            g_stage.alloc((size_t)G::P * G::TB + G::FRAG_DOUBLES + 2); g_stage.zero(stream);
            g_delta.alloc(nch * 2 * G::DMAX); g_delta.zero(stream);
            g_delta_tag.alloc(nch * 2);
            APT_CUDA(cudaMemsetAsync(g_delta_tag.p, 0xff, nch * 2 * sizeof(long long), stream));
            g_pend.alloc(nch * APT_PEND_DOUBLES); g_pend.zero(stream);
            g_scratch.alloc(nch * 2 * G::DMAX); g_scratch.zero(stream);
        } else {
            if (SH::SM_BYTES > 48 * 1024)
                APT_CUDA(cudaFuncSetAttribute(ladder_kernel<G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SH::SM_BYTES));
            // the resident ladders per SM are bounded by shared memory: ask for the largest carve-out
            APT_CUDA(cudaFuncSetAttribute(ladder_kernel<G>, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared));
            if (getenv("APTB200_DEBUG_TIMING")) {
                int occ = 0;
                cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, ladder_kernel<G>, SH::CTA, SH::SM_BYTES);
                cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, ladder_kernel<G>);
                fprintf(stderr, "[aptb200] ladder_kernel: %d threads/CTA, %zu B smem/CTA, %d regs, %d CTAs/SM (%d ladders/SM)\n", SH::CTA, (size_t)SH::SM_BYTES, fa.numRegs, occ, occ * SH::TEAMS);
            }
        }
        APT_CUDA(cudaStreamSynchronize(stream));
    }

    ~Engine() {
        if (ev0) cudaEventDestroy(ev0);
        for (auto e : evp) if (e) cudaEventDestroy(e);
        if (ev1) cudaEventDestroy(ev1);
        if (stream) cudaStreamDestroy(stream);
        if (copy_stream) cudaStreamDestroy(copy_stream);
    }

    void set_model_theta(const double* th, int ladder) {
        APT_CUDA(cudaSetDevice(device));
        if (ladder >= L) throw ApiError{"ladder index out of range"};
        APT_CUDA(cudaMemcpyAsync(model_theta.p + (size_t)std::max(ladder, 0) * P, th, P * sizeof(double), cudaMemcpyHostToDevice, stream));
        if (ladder < 0 && L > 1) {
            const size_t total = (size_t)L * P;
            replicate_row0_kernel<<<(unsigned)std::min<size_t>((total + 255) / 256, 1184), 256, 0, stream>>>(model_theta.p, P, total);
            APT_CUDA(cudaGetLastError());
        }
        APT_CUDA(cudaStreamSynchronize(stream));
    }
    void set_array(int idx, const double* v, int64_t n) {
        APT_CUDA(cudaSetDevice(device));
        if (idx < 0 || idx >= G::NARR || (size_t)n != arrays[idx].n) throw ApiError{"set_array: bad index or length"};
        APT_CUDA(cudaMemcpyAsync(arrays[idx].p, v, n * sizeof(double), cudaMemcpyHostToDevice, stream));
        refresh_derived();
    }

    static void grow_trace(DevBuf<double>& b, int L, long long oldcap, long long newcap, long long keep_rows, int cols,
                           cudaStream_t s) {
        if (cols == 0) return;
        DevBuf<double> nb;
        nb.alloc((size_t)L * newcap * cols);
        if (b.p && keep_rows > 0)
            copy_rows_2d(nb.p, (size_t)newcap * cols, b.p, (size_t)oldcap * cols, (size_t)keep_rows * cols, L, cudaMemcpyDeviceToDevice, s);
        APT_CUDA(cudaStreamSynchronize(s));
        std::swap(b.p, nb.p);
        std::swap(b.n, nb.n);
    }

    // whole-model per-declaration logProbs of `nst` states on the device (initializeModel / calculate())
    void eval_lpd(const double* d_theta, int nst, double* d_out) {
        if constexpr (G::ENGINE == 2) {
            constexpr int NBLK = 128, THREADS = 256;
            DevBuf<double> part;
            part.alloc((size_t)nst * NBLK * G::ND);
            for (int s0 = 0; s0 < nst; s0 += 32768) {
                const int n = std::min(32768, nst - s0);
                logprob_wide_kernel<G, NBLK, THREADS><<<dim3(NBLK, n), THREADS, 0, stream>>>(D, d_theta + (size_t)s0 * P, part.p + (size_t)s0 * NBLK * G::ND);
            }
            APT_CUDA(cudaGetLastError());
            logprob_wide_reduce<G, NBLK><<<(nst * G::ND + 127) / 128, 128, 0, stream>>>(part.p, nst, d_out);
            APT_CUDA(cudaGetLastError());
            APT_CUDA(cudaStreamSynchronize(stream));
            timing.launches += 2;
        } else {
            const int threads = 128;
            const int blocks = (int)(((size_t)nst * G::W + threads - 1) / threads);
            logprob_kernel<G><<<blocks, threads, 0, stream>>>(D, d_theta, nst, d_out);
            APT_CUDA(cudaGetLastError());
            timing.launches++;
        }
    }
    void refresh_lpd() { eval_lpd(theta.p, L * K, lpd.p); }

    void run(const aptmod_run_args* ra) {
        APT_CUDA(cudaSetDevice(device));
        if (ra->niter < 0) throw ApiError{"cannot specify niter < 0"};  // APT:350
        int thin = ra->thin != -1 ? ra->thin : thin_conf;               // APT:351-353
        int thin2 = ra->thin2 != -1 ? ra->thin2 : thin2_conf;
        if (thin < 1 || thin2 < 1) throw ApiError{"thin and thin2 must be >= 1"};
        last_thin = thin;
        bool reset = ra->reset != 0, resetTempering = ra->resetTempering != 0;
        if (resetAnyway) {  // APT:358-363
            reset = true; resetTempering = true; resetAnyway = false; totalIters = 1;
        }
        if (resetTempering) totalIters = 1;  // APT:365-367
        timing = aptmod_timing{};
        accswap.zero(stream);                // APT:368
        long long off, off2;
        if (reset) {                         // APT:369-377
            reset_kernel<G><<<std::min<size_t>((size_t)L * K, 65535), 64, 0, stream>>>(L, model_theta.p, theta.p, slot.p, scale.p,
                                                                                     cnt.p, blk.p, scale0.p, blk0.p, !blk_initialised);
            blk_initialised = true;
            APT_CUDA(cudaGetLastError());
            timing.launches++;
            off = 0; off2 = 0;
        } else {                             // APT:378-385
            off = rows; off2 = rows2;
            if (ra->resetMV) { off = 0; off2 = 0; }
        }
        const long long nr = ra->niter / thin, nr2 = ra->niter / thin2;
        // a continuation (reset = FALSE, APT:378-385) grows the capacity geometrically, so that a sequence of appending
        // runs does not copy the whole trace every time; the Tmax traces exist only with monitorTmax (APT:314-317)
        if (off + nr > cap) {
            long long nc = off + nr;
            if (off > 0) nc = std::max(nc, cap + cap / 2);
            grow_trace(samples, L, cap, nc, off, G::NMON, stream);
            if (monitor_tmax) grow_trace(samplesT, L, cap, nc, off, G::NMON, stream);
            grow_trace(logprobs, L, cap, nc, off, 1, stream);
            cap = nc;
        }
        if (off2 + nr2 > cap2) {
            long long nc = off2 + nr2;
            if (off2 > 0) nc = std::max(nc, cap2 + cap2 / 2);
            grow_trace(samples2, L, cap2, nc, off2, G::NMON2, stream);
            if (monitor_tmax) grow_trace(samples2T, L, cap2, nc, off2, G::NMON2, stream);
            cap2 = nc;
        }
        if (nr > ttcap) { temptraj.alloc((size_t)L * nr * K); ttcap = nr; }
        // the row counts follow the iterations that have actually run: an interrupted run (should_abort) or a failed
        // launch must not report rows that were never written
        rows = off; rows2 = off2; ttrows = 0;
        auto rows_after = [&](long long it_done) { rows = off + it_done / thin; rows2 = off2 + it_done / thin2; ttrows = it_done / thin; };
        chunk_cb = (ra->chunk_cb && nr > 0 && G::NMON > 0) ? ra->chunk_cb : nullptr;
        chunk_arg = ra->chunk_arg;
        so_ptr = (!chunk_cb && ra->samples_out && nr > 0 && G::NMON > 0) ? ra->samples_out : nullptr;
        if (so_ptr) {
            if ((long long)L * nr * G::NMON > ra->samples_out_cap) throw ApiError{"samples_out: buffer too small for n_ladders x (niter / thin) x n_monitors doubles"};
            if (!copy_stream) APT_CUDA(cudaStreamCreateWithFlags(&copy_stream, cudaStreamNonBlocking));
        }
        if (so_ptr || chunk_cb) { so_nr = nr; so_off = off; so_thin = thin; }
        if (chunk_cb && pack.n < (size_t)L * nr * G::NMON) pack.alloc((size_t)L * nr * G::NMON);
        // with a destination the run is split into at least 4 launches (8 for long runs), so that moving a piece overlaps
        // the sampling of the next one and only the last piece is left when the run ends
        const long long so_pieces = ra->niter >= 8 * 64 * (long long)thin ? 8 : 4;
        const long long so_chunk = (so_ptr || chunk_cb) ? std::max<long long>(thin, ((ra->niter + so_pieces - 1) / so_pieces + thin - 1) / thin * thin) : 0;
        refresh_lpd();
        // test hooks
        const size_t nz = (size_t)ra->niter * L * K * G::NU;
        if (ra->tabZ) { tabZ.alloc(nz * G::DMAX); APT_CUDA(cudaMemcpyAsync(tabZ.p, ra->tabZ, nz * G::DMAX * 8, cudaMemcpyHostToDevice, stream)); }
        if (ra->tabU) { tabU.alloc(nz); APT_CUDA(cudaMemcpyAsync(tabU.p, ra->tabU, nz * 8, cudaMemcpyHostToDevice, stream)); }
        if (ra->tabS) { tabS.alloc((size_t)ra->niter * L * K * 2); APT_CUDA(cudaMemcpyAsync(tabS.p, ra->tabS, (size_t)ra->niter * L * K * 16, cudaMemcpyHostToDevice, stream)); }
        if (record) {
            rec.alloc(std::max<size_t>(nz, 1)); rec.zero(stream);
            recswap.alloc(std::max<size_t>((size_t)ra->niter * L * K, 1)); recswap.zero(stream);
            rec_iters = ra->niter;
        }
        KArgs a{};
        a.D = D; a.L = L; a.ladder0 = ladder0; a.seed = seed;
        a.adapt_temps = ra->adaptTemps; a.tune1 = ra->tuneTemper1; a.tune2 = ra->tuneTemper2; a.ULT = ULT;
        a.thin = thin; a.thin2 = thin2; a.monitor_tmax = monitor_tmax;
        a.off = off; a.off2 = off2; a.cap = cap; a.cap2 = cap2; a.ttcap = ttcap;
        a.theta = theta.p; a.lpd = lpd.p; a.temps = temps.p; a.lp = lp.p; a.scale = scale.p; a.blk = blk.p; a.empir = empir.p; a.famdelta = famdelta.p; a.wpart = wpart.p;
        a.slot = slot.p; a.cnt = cnt.p; a.accswap = accswap.p;
        a.samples = samples.p; a.samples2 = samples2.p; a.samplesT = samplesT.p; a.samples2T = samples2T.p;
        a.logprobs = logprobs.p; a.temptraj = temptraj.p;
        a.tabZ = ra->tabZ ? tabZ.p : nullptr; a.tabU = ra->tabU ? tabU.p : nullptr; a.tabS = ra->tabS ? tabS.p : nullptr;
        a.rec = record ? rec.p : nullptr; a.recswap = record ? recswap.p : nullptr;
        bool aborted = false;
        if constexpr (G::ENGINE == 2) {
            long long done = 0;
            try { aborted = run_grid(a, ra, done); } catch (...) { rows_after(done); throw; }
            rows_after(done);
        } else {
            long long chunk = ra->iters_per_launch > 0 ? ra->iters_per_launch : 4096;
            if (so_chunk > 0 && ra->iters_per_launch <= 0) chunk = std::min(chunk, so_chunk);
            const int grid = (L + SH::TEAMS - 1) / SH::TEAMS;
            // Two pieces in flight: piece i + 1 is queued before the host waits for piece i, so the device never idles between
            // two pieces (the host's wait + launch was ~15 us per piece: 10 % of a 1-ms run), and an interrupt is still noticed
            // within two pieces.  The copy of a piece's rows waits for the piece's event on the second stream.
            if (!evp[0]) for (auto& e : evp) APT_CUDA(cudaEventCreate(&e));
            long long prev_it1 = -1;
            int npiece = 0;
            auto finish_piece = [&](int pi, long long it1p) {
                APT_CUDA(cudaEventSynchronize(evp[2 * pi + 1]));
                float ms = 0;
                APT_CUDA(cudaEventElapsedTime(&ms, evp[2 * pi], evp[2 * pi + 1]));
                timing.kernel_ms += ms; timing.main_kernel_ms += ms; timing.launches++; timing.main_launches++;
                rows_after(it1p);
            };
            for (long long it0 = 0; it0 < ra->niter; it0 += chunk) {
                if (ra->should_abort && ra->should_abort(ra->abort_arg)) { aborted = true; break; }
                a.it0 = it0; a.it1 = std::min<long long>(ra->niter, it0 + chunk);
                a.rng_iter0 = rngIter; a.total_iters0 = totalIters;
                const int pi = npiece & 1;
                APT_CUDA(cudaEventRecord(evp[2 * pi], stream));
                if constexpr (G::CLUSTER > 1) {
                    cudaLaunchConfig_t cfg{};
                    cfg.gridDim = dim3((unsigned)(L * G::CLUSTER)); cfg.blockDim = dim3(SH::CTA);
                    cfg.dynamicSmemBytes = SH::SM_BYTES; cfg.stream = stream;
                    cudaLaunchAttribute at[1];
                    at[0].id = cudaLaunchAttributeClusterDimension;
                    at[0].val.clusterDim.x = G::CLUSTER; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
                    cfg.attrs = at; cfg.numAttrs = 1;
                    APT_CUDA(cudaLaunchKernelEx(&cfg, ladder_kernel<G>, a));
                } else
                ladder_kernel<G><<<grid, SH::CTA, SH::SM_BYTES, stream>>>(a);
                APT_CUDA(cudaGetLastError());
                APT_CUDA(cudaEventRecord(evp[2 * pi + 1], stream));
                rngIter += (uint64_t)(a.it1 - a.it0);
                totalIters += (a.it1 - a.it0);
                if (so_ptr) APT_CUDA(cudaStreamWaitEvent(copy_stream, evp[2 * pi + 1], 0));
                copy_rows_async(a.it0, a.it1);
                if (prev_it1 >= 0) finish_piece(pi ^ 1, prev_it1);
                prev_it1 = a.it1;
                npiece++;
            }
            if (prev_it1 >= 0) finish_piece((npiece - 1) & 1, prev_it1);
        }
        // APT:548: the model is left parameterised with the T=1 state
        model_from_rung1_kernel<G><<<std::min(L, 65535), 64, 0, stream>>>(L, theta.p, slot.p, model_theta.p);
        APT_CUDA(cudaGetLastError());
        timing.launches++;
        APT_CUDA(cudaStreamSynchronize(stream));
        if (so_ptr) { so_ptr = nullptr; APT_CUDA(cudaStreamSynchronize(copy_stream)); }
        chunk_cb = nullptr;
        launches_total += timing.launches;
        if (aborted) throw ApiError{"run interrupted by the caller"};
    }

    // grid engine: one launch per (iteration, stage, pass); chunks of iterations are self-contained (prologue
    // launch proposes, the last launch of the chunk does not propose ahead) so that a run can be interrupted
    // between chunks with the state consistent.  Returns true when interrupted.
    bool run_grid(KArgs& a, const aptmod_run_args* ra, long long& it_done) {
        if constexpr (G::ENGINE == 2) {
            GArgs g{};
            const int nchain = L * K;
            const int npass = (nchain + G::TB - 1) / G::TB;
            size_t smem = GridShape<G>::smem_bytes(2);
            if (grid_occ_known == 0 && smem > 48 * 1024) G::set_stage_smem(smem);
            if (grid_occ_known == 0) {
                grid_occ = std::max(1, G::stage_occupancy(0, smem));
                grid_occ_known = 1;
            }
            const int grid = num_sms * std::min(grid_occ, 8);
            long long chunk = ra->iters_per_launch > 0 ? ra->iters_per_launch : 256;
            const uint64_t rng0 = rngIter;
            const long long tot0 = totalIters;
            for (long long it0 = 0; it0 < ra->niter; it0 += chunk) {
                if (ra->should_abort && ra->should_abort(ra->abort_arg)) return true;
                const long long it1 = std::min<long long>(ra->niter, it0 + chunk);
                // iterations are addressed by their index within the RUN (thinning, swap direction, random slots,
                // injected tables all depend on it); the counters at run start are the bases
                a.rng_iter0 = rng0; a.total_iters0 = tot0;
                a.it0 = it0; a.it1 = it1;
                g.k = a;
                g.partial = g_partial.p; g.ticket = g_ticket.p; g.pend = g_pend.p; g.scratch = g_scratch.p;
                g.npass = npass;
                g.dbg = getenv("APTB200_DEBUG_TIMING") ? g_dbg.p : nullptr;
                g.delta = getenv("APTB200_NO_DELTA_CTA") ? nullptr : g_delta.p; g.delta_tag = g_delta_tag.p;
                APT_CUDA(cudaEventRecord(ev0, stream));
                g.flags = g_flags.p; g.seq0 = launch_seq; g.stage_img = g_stage.p;
                const long long nl = (it1 - it0) * G::NSTAGE * npass;  // steps (passes over the data) of this launch
                launch_seq += (unsigned long long)nl + 1;  // the proposals of the first step are published as seq0 + 1, step n as seq0 + n + 2
                if (grid < 2) throw ApiError{"grid engine: the device must hold at least 2 resident CTAs (a control CTA and a streaming CTA)"};
                APT_CUDA(G::launch_run(g, grid, smem, stream));
                APT_CUDA(cudaGetLastError());
                APT_CUDA(cudaEventRecord(ev1, stream));
                APT_CUDA(cudaEventSynchronize(ev1));
                float ms = 0;
                APT_CUDA(cudaEventElapsedTime(&ms, ev0, ev1));
                timing.kernel_ms += ms; timing.main_kernel_ms += ms; timing.launches += 1; timing.main_launches += nl;  // main_launches: passes over the data inside the persistent launch
                rngIter += (uint64_t)(it1 - it0);
                totalIters += (it1 - it0);
                it_done = it1;
                copy_rows_async(it0, it1);
                if (getenv("APTB200_DEBUG_TIMING")) {
                    auto v = d2h(g_dbg.p, 80);
                    fprintf(stderr, "[aptb200] control CTA cycles (last step of the launch): delta %lld | preload %lld | wait for the streaming CTAs %lld | reduce %lld | finish %lld | ladder (to the swaps) %lld | propose+publish %lld | deferred %lld\n", v[7], v[6] - v[5] - v[7], v[0] - v[6], v[1] - v[0], v[2] - v[1], v[3] - v[2], v[4] - v[3], v[15] - v[4]);
                    {   // hand-over of the last step: its proposals were published by the step before (same parity slot)
                        const int par = (int)((launch_seq - 1) & 1);  // seq of the last step = seq0 + nl = launch_seq - 1
                        fprintf(stderr, "   hand-over (CTA 0, ns): published -> seen %lld | seen -> parameters staged %lld\n", v[22 + par] - v[20 + par], v[24 + par] - v[22 + par]);
                    }
                    fprintf(stderr, "   ladder step: load %lld | lp+philox+log %lld | K^2 exp %lld | rowsum %lld | normalise+log %lld | chain %lld\n", v[8] - v[2], v[9] - v[8], v[10] - v[9], v[11] - v[10], v[12] - v[11], v[13] - v[12]);
                }
            }
        }
        return false;
    }
    int grid_occ_known = 0;
    long long measure_fp64() const {
        cudaSetDevice(device);
        int sms = 148;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
        const int grid = sms * 2, iters = 4000;
        DevBuf<double> o;
        o.alloc((size_t)grid * 256);
        fp64_peak_kernel<<<grid, 256, 0, stream>>>(o.p, iters, 0.999, 1e-3);  // warm-up
        float best = 1e30f;
        for (int rep = 0; rep < 3; rep++) {
            cudaEventRecord(ev0, stream);
            fp64_peak_kernel<<<grid, 256, 0, stream>>>(o.p, iters, 0.999, 1e-3);
            cudaEventRecord(ev1, stream);
            cudaEventSynchronize(ev1);
            float ms = 0;
            cudaEventElapsedTime(&ms, ev0, ev1);
            best = std::min(best, ms);
        }
        const double flops = 2.0 * grid * 256.0 * iters * 256.0;
        return (long long)(flops / (best * 1e-3) / 1e9);
    }
    // derived arrays live behind the model's own arrays in Data::a[]
    std::vector<DevBuf<double>> derived;
    void refresh_derived() {
        if constexpr (G::NDER > 0) {
            static_assert(G::NARR + G::NDER <= APT_MAX_ARR, "too many arrays");
            if (derived.empty()) {
                derived.resize(G::NDER);
                for (int k = 0; k < G::NDER; k++) {
                    derived[k].alloc((size_t)std::max<long>(G::der_inst(k) * G::der_width(k), 1));
                    D.a[G::NARR + k] = derived[k].p;
                }
            }
            for (int k = 0; k < G::NDER; k++) {
                const long n = G::der_inst(k);
                if (n > 0) derived_kernel<G><<<(unsigned)std::min<long>((n + 255) / 256, 4096), 256, 0, stream>>>(D, k, n, derived[k].p);
            }
            APT_CUDA(cudaGetLastError());
        }
    }

    // whole-model logProb of caller-provided states (host pointers): out_lpd [n][ND]
    void logprob(const double* th, int n, double* out_lpd) {
        APT_CUDA(cudaSetDevice(device));
        DevBuf<double> dth, dout;
        dth.alloc((size_t)n * P); dout.alloc((size_t)n * G::ND);
        APT_CUDA(cudaMemcpyAsync(dth.p, th, (size_t)n * P * 8, cudaMemcpyHostToDevice, stream));
        eval_lpd(dth.p, n, dout.p);
        APT_CUDA(cudaMemcpyAsync(out_lpd, dout.p, (size_t)n * G::ND * 8, cudaMemcpyDeviceToHost, stream));
        APT_CUDA(cudaStreamSynchronize(stream));
    }

    // calculateWAIC (APT:593-635): out3 = {WAIC, logAvgProb, pWAIC}
    void waic(long long nburnin, int ladder, double* out3) {
        APT_CUDA(cudaSetDevice(device));
        if (ladder >= L) throw ApiError{"ladder index out of range"};
        if constexpr (G::NDATA == 0) throw ApiError{"WAIC cannot be calculated, as no data nodes were detected in the model."};  // APT:331
        else if constexpr ((size_t)WaicShape<G>::S * P * 8 > 48 * 1024) throw ApiError{"calculateWAIC: too many parameters for the shared-memory staging of the device kernel"};
        else {
            if (nburnin < 0) nburnin = 0;
            const long long nb = (nburnin + last_thin - 1) / last_thin;  // nburninPostThinning, APT:605
            const long long n = rows - nb;                              // numMCMCSamples, APT:606
            out3[1] = out3[2] = std::nan("");
            if (n < 2) { out3[0] = -INFINITY; return; }               // APT:607-610
            DevBuf<double> res;
            res.alloc((size_t)2 * G::NDATA);
            const unsigned grid = (unsigned)((G::NDATA + WAIC_THREADS - 1) / WAIC_THREADS);
            APT_CUDA(cudaEventRecord(ev0, stream));
            waic_kernel<G><<<grid, WAIC_THREADS, 0, stream>>>(D, samples.p, cap, nb, n, ladder < 0 ? 0 : ladder, ladder < 0 ? L : 1,
                                                            model_theta.p, res.p);
            APT_CUDA(cudaGetLastError());
            APT_CUDA(cudaEventRecord(ev1, stream));
            auto v = d2h(res.p, (size_t)2 * G::NDATA);
            float ms = 0;
            APT_CUDA(cudaEventElapsedTime(&ms, ev0, ev1));
            waic_ms = ms;
            launches_total++;
            double lavg = 0, pw = 0;  // APT:623-629: summed in data-node order
            for (long j = 0; j < G::NDATA; j++) { lavg += v[(size_t)j]; pw += v[(size_t)G::NDATA + j]; }
            out3[0] = -2 * (lavg - pw); out3[1] = lavg; out3[2] = pw;   // APT:630
        }
    }
    double waic_ms = 0;

    // device-resident, compact [L][rows][cols] view of a trace for collectives (NCCL all-gather); no host staging
    DevBuf<double> compact;
    void device_ptr(int what, void** ptr, int64_t* n, void** strm) {
        APT_CUDA(cudaSetDevice(device));
        const DevBuf<double>* b; long long capr, nrows; int cols;
        switch (what) {
            case APTMOD_SAMPLES: b = &samples; capr = cap; nrows = rows; cols = G::NMON; break;
            case APTMOD_SAMPLES_TMAX: b = &samplesT; capr = cap; nrows = rows; cols = G::NMON; break;
            case APTMOD_SAMPLES2: b = &samples2; capr = cap2; nrows = rows2; cols = G::NMON2; break;
            case APTMOD_LOGPROBS: b = &logprobs; capr = cap; nrows = rows; cols = 1; break;
            case APTMOD_TEMPTRAJ: b = &temptraj; capr = ttcap; nrows = ttrows; cols = K; break;
            default: throw ApiError{"this item is not available as a device trace"};
        }
        *n = (int64_t)L * nrows * cols;
        *strm = (void*)stream;
        if (capr == nrows) { *ptr = b->p; return; }
        if (compact.n < (size_t)std::max<int64_t>(*n, 1)) compact.alloc((size_t)std::max<int64_t>(*n, 1));
        if (*n) copy_rows_2d(compact.p, (size_t)nrows * cols, b->p, (size_t)capr * cols, (size_t)nrows * cols, L, cudaMemcpyDeviceToDevice, stream);
        *ptr = compact.p;
    }

    long long info(int what) const {
        switch (what) {
            case APTMOD_INFO_P: return P;
            case APTMOD_INFO_K: return K;
            case APTMOD_INFO_NU: return G::NU;
            case APTMOD_INFO_ND: return G::ND;
            case APTMOD_INFO_NMON: return G::NMON;
            case APTMOD_INFO_NMON2: return G::NMON2;
            case APTMOD_INFO_DMAX: return G::DMAX;
            case APTMOD_INFO_ROWS: return rows;
            case APTMOD_INFO_ROWS2: return rows2;
            case APTMOD_INFO_TTROWS: return ttrows;
            case APTMOD_INFO_L: return L;
            case APTMOD_INFO_LAUNCHES: return launches_total;
            case APTMOD_INFO_TOTALITERS: return totalIters;
            case APTMOD_INFO_W: return G::W;
            case APTMOD_INFO_CTA: return SH::CTA;
            case APTMOD_INFO_ENGINE: return G::ENGINE;
            case APTMOD_INFO_BLKSZ: return G::BLKSZ;
            case APTMOD_INFO_FP64_GFLOPS: return measure_fp64();
        }
        return -1;
    }

    template <class T>
    std::vector<T> d2h(const T* p, size_t n) {
        std::vector<T> v(n);
        if (n) APT_CUDA(cudaMemcpyAsync(v.data(), p, n * sizeof(T), cudaMemcpyDeviceToHost, stream));
        APT_CUDA(cudaStreamSynchronize(stream));
        return v;
    }
    void trace_out(const DevBuf<double>& b, long long capr, long long nrows, int cols, int ladder, double* out, int64_t capn,
                   int64_t& n) {
        if (ladder >= L) throw ApiError{"ladder index out of range"};
        const int nl = ladder < 0 ? L : 1;
        n = (int64_t)nl * nrows * cols;
        if (n > capn) throw ApiError{"output buffer too small"};
        if (n == 0) return;
        const double* src = b.p + (size_t)(ladder < 0 ? 0 : ladder) * capr * cols;
        if (!b.p) throw ApiError{"this trace is not kept (monitorTmax = FALSE)"};
        copy_rows_2d(out, (size_t)nrows * cols, src, (size_t)capr * cols, (size_t)nrows * cols, nl, cudaMemcpyDeviceToHost, stream);
        APT_CUDA(cudaStreamSynchronize(stream));
    }

    // One ladder's matrix in COLUMN-major order (what R's REAL() of a matrix is): traces are transposed on the device and
    // copied straight into the caller's buffer; the small per-rung matrices are transposed on the host.
    DevBuf<double> transposed;
    int64_t get_cm(int what, int ladder, double* out, int64_t capn) {
        APT_CUDA(cudaSetDevice(device));
        if (ladder < 0 || ladder >= L) throw ApiError{"column-major output is per ladder: ladder index out of range"};
        const DevBuf<double>* b = nullptr; long long capr = 0, nrows = 0; int cols = 0;
        switch (what) {
            case APTMOD_SAMPLES: b = &samples; capr = cap; nrows = rows; cols = G::NMON; break;
            case APTMOD_SAMPLES_TMAX: b = &samplesT; capr = cap; nrows = rows; cols = G::NMON; break;
            case APTMOD_SAMPLES2: b = &samples2; capr = cap2; nrows = rows2; cols = G::NMON2; break;
            case APTMOD_SAMPLES2_TMAX: b = &samples2T; capr = cap2; nrows = rows2; cols = G::NMON2; break;
            case APTMOD_TEMPTRAJ: b = &temptraj; capr = ttcap; nrows = ttrows; cols = K; break;
            default: break;
        }
        if (b) {
            const int64_t n = (int64_t)nrows * cols;
            if (n > capn) throw ApiError{"output buffer too small"};
            if (n == 0) return 0;
            if (!b->p) throw ApiError{"this trace is not kept (monitorTmax = FALSE)"};
            if (transposed.n < (size_t)n) transposed.alloc((size_t)n);
            const dim3 grid((unsigned)((nrows + 31) / 32), (unsigned)((cols + 31) / 32));
            transpose_kernel<<<grid, dim3(32, 8), 0, stream>>>(b->p + (size_t)ladder * capr * cols, (long)nrows, cols, transposed.p);
            APT_CUDA(cudaGetLastError());
            APT_CUDA(cudaMemcpyAsync(out, transposed.p, (size_t)n * 8, cudaMemcpyDeviceToHost, stream));
            APT_CUDA(cudaStreamSynchronize(stream));
            launches_total++;
            return n;
        }
        int r = 1, c = 1;
        switch (what) {
            case APTMOD_ACCSWAP: r = K; c = K; break;
            case APTMOD_THETA: r = K; c = P; break;
            case APTMOD_SCALE: r = K; c = G::NU; break;
            case APTMOD_PROPCOV: r = K; c = G::BLKSZ; break;
            default: return get(what, ladder, out, capn);  // vectors: the same in either order
        }
        std::vector<double> tmp((size_t)std::max(r * c, 1));
        const int64_t n = get(what, ladder, tmp.data(), (int64_t)tmp.size());
        if (n > capn) throw ApiError{"output buffer too small"};
        for (int i = 0; i < r; i++)
            for (int j = 0; j < c; j++) out[(size_t)j * r + i] = tmp[(size_t)i * c + j];
        return n;
    }

    // copies into caller-allocated host memory; returns the number of doubles written
    int64_t get(int what, int ladder, double* out, int64_t capn) {
        APT_CUDA(cudaSetDevice(device));
        int64_t n = 0;
        const int l = ladder < 0 ? 0 : ladder;
        if (ladder >= L) throw ApiError{"ladder index out of range"};
        switch (what) {
            case APTMOD_SAMPLES: trace_out(samples, cap, rows, G::NMON, ladder, out, capn, n); return n;
            case APTMOD_SAMPLES_TMAX: trace_out(samplesT, cap, rows, G::NMON, ladder, out, capn, n); return n;
            case APTMOD_SAMPLES2: trace_out(samples2, cap2, rows2, G::NMON2, ladder, out, capn, n); return n;
            case APTMOD_SAMPLES2_TMAX: trace_out(samples2T, cap2, rows2, G::NMON2, ladder, out, capn, n); return n;
            case APTMOD_LOGPROBS: trace_out(logprobs, cap, rows, 1, ladder, out, capn, n); return n;
            case APTMOD_TEMPTRAJ: trace_out(temptraj, ttcap, ttrows, K, ladder, out, capn, n); return n;
            case APTMOD_TEMPS:
            case APTMOD_LOGPROBTEMPS: {
                const int nl = ladder < 0 ? L : 1;
                n = (int64_t)nl * K;
                if (n > capn) throw ApiError{"output buffer too small"};
                auto v = d2h((what == APTMOD_TEMPS ? temps.p : lp.p) + (size_t)l * K, (size_t)n);
                std::memcpy(out, v.data(), n * 8);
                return n;
            }
            case APTMOD_ACCSWAP: {
                n = (int64_t)K * K;
                if (n > capn) throw ApiError{"output buffer too small"};
                auto v = d2h(accswap.p + (size_t)l * K * K, (size_t)n);
                for (int64_t i = 0; i < n; i++) out[i] = v[i];
                return n;
            }
            case APTMOD_THETA: {  // [K][P] in rung order
                n = (int64_t)K * P;
                if (n > capn) throw ApiError{"output buffer too small"};
                auto v = d2h(theta.p + (size_t)l * K * P, (size_t)n);
                auto s = d2h(slot.p + (size_t)l * K, (size_t)K);
                for (int k = 0; k < K; k++) std::memcpy(out + (size_t)k * P, v.data() + (size_t)s[k] * P, P * 8);
                return n;
            }
            case APTMOD_MODEL_THETA: {
                n = P;
                if (n > capn) throw ApiError{"output buffer too small"};
                auto v = d2h(model_theta.p + (size_t)l * P, (size_t)P);
                std::memcpy(out, v.data(), P * 8);
                return n;
            }
            case APTMOD_SCALE: {
                n = (int64_t)K * G::NU;
                if (n > capn) throw ApiError{"output buffer too small"};
                auto v = d2h(scale.p + (size_t)l * K * G::NU, (size_t)n);
                std::memcpy(out, v.data(), n * 8);
                return n;
            }
            case APTMOD_PROPCOV: {
                n = (int64_t)K * G::BLKSZ;
                if (n > capn) throw ApiError{"output buffer too small"};
                auto v = d2h(blk.p + (size_t)l * K * G::BLKSZ, (size_t)n);
                std::memcpy(out, v.data(), n * 8);
                return n;
            }
            case APTMOD_DECISIONS: {  // [iters][K][NU]
                if (!record) throw ApiError{"decision recording was not enabled at build"};
                n = rec_iters * K * G::NU;
                if (n > capn) throw ApiError{"output buffer too small"};
                auto v = d2h(rec.p, (size_t)rec_iters * L * K * G::NU);
                for (long long it = 0; it < rec_iters; it++)
                    for (int i = 0; i < K * G::NU; i++) out[it * K * G::NU + i] = v[((size_t)it * L + l) * K * G::NU + i];
                return n;
            }
            case APTMOD_SWAPREC: {  // [iters][K]  value = iTo | accepted << 16
                if (!record) throw ApiError{"decision recording was not enabled at build"};
                n = rec_iters * K;
                if (n > capn) throw ApiError{"output buffer too small"};
                auto v = d2h(recswap.p, (size_t)rec_iters * L * K);
                for (long long it = 0; it < rec_iters; it++)
                    for (int i = 0; i < K; i++) out[it * K + i] = v[((size_t)it * L + l) * K + i];
                return n;
            }
        }
        throw ApiError{"unknown getter"};
    }
};

inline void set_err(char* err, int errlen, const std::string& m) {
    if (err && errlen > 0) {
        std::snprintf(err, (size_t)errlen, "%s", m.c_str());
    }
}

}  // namespace apt

#define APT_GUARD(body)                                                    \
    try {                                                                  \
        body;                                                              \
        return 0;                                                          \
    } catch (const apt::ApiError& e) {                                     \
        apt::set_err(err, errlen, e.msg);                                  \
        return 1;                                                          \
    } catch (const std::exception& e) {                                    \
        apt::set_err(err, errlen, e.what());                               \
        return 1;                                                          \
    } catch (...) {                                                        \
        apt::set_err(err, errlen, "unknown error");                        \
        return 1;                                                          \
    }

#define APT_EXPORT __attribute__((visibility("default")))
#define APT_DEFINE_MODULE(G)                                                                                        \
    extern "C" {                                                                                                    \
    APT_EXPORT int aptmod_abi_version(void) { return APTMOD_ABI_VERSION; }                                                     \
    APT_EXPORT int aptmod_create(const aptmod_init* in, void** out, char* err, int errlen) {                                   \
        *out = nullptr;                                                                                             \
        apt::Engine<G>* e = nullptr;                                                                                \
        try {                                                                                                       \
            e = new apt::Engine<G>();                                                                               \
            e->create(in);                                                                                          \
            *out = e;                                                                                               \
            return 0;                                                                                               \
        } catch (const apt::ApiError& x) {                                                                          \
            apt::set_err(err, errlen, x.msg);                                                                       \
        } catch (const std::exception& x) {                                                                         \
            apt::set_err(err, errlen, x.what());                                                                    \
        } catch (...) {                                                                                             \
            apt::set_err(err, errlen, "unknown error");                                                             \
        }                                                                                                           \
        delete e;                                                                                                   \
        return 1;                                                                                                   \
    }                                                                                                               \
    APT_EXPORT void aptmod_destroy(void* h) { delete static_cast<apt::Engine<G>*>(h); }                                        \
    APT_EXPORT int aptmod_run(void* h, const aptmod_run_args* ra, char* err, int errlen) {                                     \
        APT_GUARD(static_cast<apt::Engine<G>*>(h)->run(ra))                                                         \
    }                                                                                                               \
    APT_EXPORT int aptmod_get(void* h, int what, int ladder, double* out, int64_t cap, int64_t* n, char* err, int errlen) {    \
        APT_GUARD(*n = static_cast<apt::Engine<G>*>(h)->get(what, ladder, out, cap))                                \
    }                                                                                                               \
    APT_EXPORT int aptmod_get_cm(void* h, int what, int ladder, double* out, int64_t cap, int64_t* n, char* err, int errlen) { \
        APT_GUARD(*n = static_cast<apt::Engine<G>*>(h)->get_cm(what, ladder, out, cap))                             \
    }                                                                                                               \
    APT_EXPORT long long aptmod_info(void* h, int what) { return static_cast<apt::Engine<G>*>(h)->info(what); }                \
    APT_EXPORT int aptmod_logprob(void* h, const double* th, int n, double* out, char* err, int errlen) {                      \
        APT_GUARD(static_cast<apt::Engine<G>*>(h)->logprob(th, n, out))                                             \
    }                                                                                                               \
    APT_EXPORT int aptmod_set_model_theta(void* h, const double* th, int ladder, char* err, int errlen) {                      \
        APT_GUARD(static_cast<apt::Engine<G>*>(h)->set_model_theta(th, ladder))                                     \
    }                                                                                                               \
    APT_EXPORT int aptmod_set_array(void* h, int idx, const double* v, int64_t n, char* err, int errlen) {                     \
        APT_GUARD(static_cast<apt::Engine<G>*>(h)->set_array(idx, v, n))                                            \
    }                                                                                                               \
    APT_EXPORT int aptmod_waic(void* h, long long nburnin, int ladder, double* out3, char* err, int errlen) {                  \
        APT_GUARD(static_cast<apt::Engine<G>*>(h)->waic(nburnin, ladder, out3))                                     \
    }                                                                                                               \
    APT_EXPORT double aptmod_waic_ms(void* h) { return static_cast<apt::Engine<G>*>(h)->waic_ms; }                             \
    APT_EXPORT void aptmod_timing_get(void* h, aptmod_timing* t) { *t = static_cast<apt::Engine<G>*>(h)->timing; }             \
    APT_EXPORT void* aptmod_stream(void* h) { return (void*)static_cast<apt::Engine<G>*>(h)->stream; }                          \
    APT_EXPORT int aptmod_device_ptr(void* h, int what, void** ptr, int64_t* n, void** strm) {                                 \
        try { static_cast<apt::Engine<G>*>(h)->device_ptr(what, ptr, n, strm); return 0; } catch (...) { return 1; }            \
    }                                                                                                               \
    }
