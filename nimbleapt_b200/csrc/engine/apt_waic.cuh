// apt_waic.cuh — calculateWAIC (/root/reference/nimbleAPT/R/APT_functions.R:593-635) on the device.
//
// The reference loops over the saved rung-1 samples, copies each into the model, recalculates the data nodes
// and fills logPredProbs[sample, dataNode] (APT:616-622); then, per data node, forms
//   maxLogPred + log(mean(exp(logPredProbs[, j] - maxLogPred)))  and  var(logPredProbs[, j])   (APT:623-629).
// Here a thread owns one data node and walks the samples twice (pass 0: max and mean, pass 1: the two sums about
// them), so the samples x dataNodes matrix is never stored: for config 3 it would be rows x 1e6 doubles.  The
// parameter vectors of a batch of samples are staged in shared memory (every thread of the CTA reads the same
// one: broadcast); the CTA's rows of the data arrays are re-read once per sample and stay in L1/L2.
// The per-node results go to HBM and are added up on the host in node order, as APT:623-629 does.
#pragma once
#include "apt_engine.cuh"

namespace apt {

constexpr int WAIC_THREADS = 128;

template <class G>
struct WaicShape {
    // samples per shared-memory batch: at most 32, at most 32 KB of parameter vectors
    static constexpr int S = G::P * 32 * 8 <= 32768 ? 32 : (32768 / (G::P * 8) > 0 ? 32768 / (G::P * 8) : 1);
};

// samples: mvSamples of the local ladders [L][cap][NMON]; sample q of the walk is row nb + q % n of ladder
// l0 + q / n (n = rows - nb post-burn-in rows per ladder, nl ladders pooled).  model_theta [L][P]: the model's
// current values, kept for every node that is not monitored (APT:614,617).  out: [2][NDATA].
template <class G>
__global__ void __launch_bounds__(WAIC_THREADS) waic_kernel(Data D, const double* __restrict__ samples, long long cap, long long nb,
                                                             long long n, int l0, int nl, const double* __restrict__ model_theta,
                                                             double* __restrict__ out) {
    constexpr int S = WaicShape<G>::S, P = G::P;
    __shared__ double th[S * P];
    const long j = (long)blockIdx.x * WAIC_THREADS + threadIdx.x;
    const bool act = j < G::NDATA;
    const long long total = n * nl;
    double mx = -CUDART_INF, sum = 0.0, mean = 0.0, se = 0.0, sv = 0.0;
    for (int pass = 0; pass < 2; ++pass) {
        for (long long q0 = 0; q0 < total; q0 += S) {
            const int ns = (int)(total - q0 < S ? total - q0 : S);
            __syncthreads();
            for (int i = threadIdx.x; i < ns * P; i += WAIC_THREADS) {
                const long long q = q0 + i / P;
                th[i] = model_theta[(size_t)(l0 + q / n) * P + i % P];
            }
            __syncthreads();
            for (int i = threadIdx.x; i < ns * G::NMON; i += WAIC_THREADS) {
                const int s = i / G::NMON, c = i % G::NMON;
                const long long q = q0 + s;
                const int code = G::mon(c);
                if (code >= 0) th[s * P + code] = samples[((size_t)(l0 + q / n) * cap + nb + q % n) * G::NMON + c];
            }
            __syncthreads();
            if (act) {
                for (int s = 0; s < ns; ++s) {
                    const double lp = G::lp_data(D, th + s * P, j);
                    if (pass == 0) {
                        mx = (lp > mx || lp != lp) ? lp : mx;
                        sum += lp;
                    } else {
                        se += exp(lp - mx);
                        const double d = lp - mean;
                        sv += d * d;
                    }
                }
            }
        }
        if (pass == 0) mean = sum / (double)total;
    }
    if (act) {
        out[j] = mx + log(se / (double)total);
        out[G::NDATA + j] = sv / (double)(total - 1);
    }
}

}  // namespace apt
