// apt_grid.cuh — the grid-wide engine: likelihoods too large for one CTA per chain (config 3: N = 1e6 rows,
// p = 50, 16 rungs; /root/reference/nimbleAPT/R/APT_functions.R:825 makes every tempered update its own serial
// pass over all N data nodes).
//
// Here ONE kernel launch per sampler stage serves the tempered update of up to TB chains (all 16 rungs):
//   1. every CTA stages the TB proposed parameter vectors in shared memory (param-major, chain-minor);
//   2. the grid streams the data once — coalesced, vectorised (double2, ld.global.cs) fp64 loads, R rows per
//      thread — and each thread accumulates the log-likelihood of its rows for all TB chains in registers
//      (inner products interchanged by the lowering so that a data element is loaded once for all chains);
//   3. warp-shuffle + shared-memory reduction per CTA, partial sums to global memory, a ticket decides which
//      CTA finishes; that CTA reduces the partials in CTA order (deterministic, no fp atomics) and then runs
//      the control step out of global memory: tempered accept/reject, Robbins-Monro adaptation, (after the
//      last stage) the swap phase, ladder adaptation and trace write, and finally the proposals of the NEXT
//      launch ("propose-ahead"), so that an iteration costs exactly one launch per stage and pass.
// Bounding roofline: HBM — 8 N (p+1) algorithmic bytes per pass.
#pragma once
#include "apt_engine.cuh"

namespace apt {

// software prefetch of streamed data one row-step ahead (L2 by default; -DAPT_PREFETCH_L1 pulls into L1)
__device__ __forceinline__ void prefetch_data(const void* p) {
#ifdef APT_PREFETCH_L1
    asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
#else
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
#endif
}
__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
// Ampere-style asynchronous global -> shared copies (LDGSTS); each thread only ever waits for its own copies
__device__ __forceinline__ void cp_async8(void* smem, const void* gmem) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(s), "l"(gmem) : "memory");
}
// 16-byte variant: bypasses L1 (.cg), the data is consumed from shared memory only
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// FP64 tensor-core step D (8 x 8) += A (8 x 4, row) * B (4 x 8, col): lane holds A[lane / 4][lane % 4], B[lane % 4][lane / 4] and
// D[lane / 4][2 (lane % 4) + {0, 1}].  Same pipe as DFMA on sm_100a, one issue slot per 256 FMAs.
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// Programmatic dependent launch: a pass kernel is launched with programmatic stream serialisation, so that its CTAs become
// resident as the previous pass's CTAs leave (all but the one that runs the control step leave right after streaming), prime
// their data rings -- the design matrix does not depend on the previous pass -- and then block in pdl_wait() until the previous
// grid has completed and its writes (accepted states, ladder, next proposals) are visible.  Removes the launch gap and the
// pipeline fill from the critical path between two passes.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

struct GArgs {
    KArgs k;
    double* partial;      // [grid][2][TB] per-CTA partial sums
    unsigned int* ticket; // arrival counter
    double* pend;         // [L*K][APT_PEND_DOUBLES] pending proposals (old log-probs, Jacobian, old scalar value)
    double* scratch;      // [L*K][2*DMAX]  normals / old block values per chain
    long long it;         // iteration (0-based within the run) this launch finishes
    int stage, pass, npass;
    int next_stage, next_pass;  // what to propose ahead (-1: nothing, the run ends or is interrupted)
    long long next_it;
    int last_of_iter;     // this launch completes the iteration: swap phase, ladder adaptation, output
    long long* dbg;       // optional phase timestamps of the finishing CTA (APTB200_DEBUG_TIMING=1)
    double* delta;        // [L*K][2*DMAX] block-proposal increments precomputed for the next launch (+ work space)
    long long* delta_tag; // [L*K][2]  (iteration, timesAdapted) the increments were computed for
};

#define APT_PEND_DOUBLES 12

template <int NG>
__device__ __forceinline__ void store_pend(double* dst, const Pend<NG>& p, int rank) {
    static_assert(NG + 2 <= APT_PEND_DOUBLES, "too many dependent groups for the grid engine");
    if (rank == 0) {
#pragma unroll
        for (int g = 0; g < NG; g++) dst[g] = p.od[g];
        dst[NG] = p.propLogScale;
        dst[NG + 1] = p.cur0;
    }
}
template <int NG>
__device__ __forceinline__ void load_pend(const double* src, Pend<NG>& p) {
#pragma unroll
    for (int g = 0; g < NG; g++) p.od[g] = src[g];
    p.propLogScale = src[NG];
    p.cur0 = src[NG + 1];
}

// context of chain q = (ladder l, rung kk) working on global memory, for control tile `tile`
template <class G>
__device__ __forceinline__ Ctx<G> grid_ctx(const GArgs& a, int l, int q, long long it, int tid) {
    constexpr int K = G::K, W = G::W;
    Ctx<G> c;
    c.rank = tid % W;
    c.tile_id = tid / W;
    c.ntiles = G::GRID_THREADS / W;
    c.tmask = (W == 32) ? 0xffffffffu : (((1u << W) - 1u) << ((tid & 31) & ~(W - 1)));
    c.theta = a.k.theta + (size_t)l * K * G::P;
    c.lpd = a.k.lpd + (size_t)l * K * G::ND;
    c.slot = a.k.slot + (size_t)l * K;
    c.T = a.k.temps + (size_t)l * K;
    c.scale = a.k.scale + (size_t)l * K * G::NU;
    c.cnt = a.k.cnt + (size_t)l * K * G::NU * 3;
    c.blk = a.k.blk + (size_t)l * K * G::BLKSZ;
    c.empir = a.k.empir + (size_t)l * K * G::EMPSZ;
    c.famdelta = nullptr;
    c.scratch = a.scratch + (size_t)q * 2 * G::DMAX;
    c.rng.seed = a.k.seed;
    c.rng.ladder = a.k.ladder0 + (uint32_t)l;
    c.rng.iter = a.k.rng_iter0 + (uint64_t)it;
    c.tabZ = a.k.tabZ ? a.k.tabZ + ((size_t)it * a.k.L + l) * K * G::NU * G::DMAX : nullptr;
    c.tabU = a.k.tabU ? a.k.tabU + ((size_t)it * a.k.L + l) * K * G::NU : nullptr;
    c.rec = a.k.rec ? a.k.rec + ((size_t)it * a.k.L + l) * K * G::NU : nullptr;
    return c;
}

template <class G>
__device__ __forceinline__ void grid_propose_ahead(const GArgs& a, int tid) {
    constexpr int K = G::K, W = G::W, TB = G::TB;
    if (a.next_stage < 0) return;
    const int tile = tid / W;
    if (tile >= TB) return;
    const int q = a.next_pass * TB + tile;
    if (q >= a.k.L * K) return;
    const int l = q / K, kk = q % K;
    Ctx<G> c = grid_ctx<G>(a, l, q, a.next_it, tid);
    const double* delta = nullptr;
    if (a.delta && G::stage_is_block(a.next_stage)) {
        // valid only if computed for this very iteration and if the rung has not adapted since (APT:859-868)
        const long long t_it = __ldcg(&a.delta_tag[(size_t)q * 2]), t_ad = __ldcg(&a.delta_tag[(size_t)q * 2 + 1]);
        const int nad = c.cnt[(kk * G::NU + G::stage_unit0(a.next_stage)) * 3 + 2];
        if (t_it == a.next_it && t_ad == (long long)nad) delta = a.delta + (size_t)q * 2 * G::DMAX;
    }
    G::propose_stage(a.next_stage, c, a.k.D, kk, a.pend + (size_t)q * APT_PEND_DOUBLES, delta);
}

// the dedicated CTA: increments of the NEXT launch's block proposals, computed while the others stream
template <class G>
__device__ __forceinline__ void grid_precompute_delta(const GArgs& a, int tid) {
    constexpr int K = G::K, W = G::W, TB = G::TB;
    const int tile = tid / W;
    if (tile >= TB) return;
    const int q = a.next_pass * TB + tile;
    if (q >= a.k.L * K) return;
    const int l = q / K, kk = q % K;
    Ctx<G> c = grid_ctx<G>(a, l, q, a.next_it, tid);
    double* out = a.delta + (size_t)q * 2 * G::DMAX;
    const int nad = c.cnt[(kk * G::NU + G::stage_unit0(a.next_stage)) * 3 + 2];
    G::delta_stage(a.next_stage, c, kk, out + G::DMAX, out);
    if (c.rank == 0) {
        a.delta_tag[(size_t)q * 2] = a.next_it;
        a.delta_tag[(size_t)q * 2 + 1] = (long long)nad;
    }
}

// first launch of a run: only the proposals of (stage 0, pass 0)
template <class G>
__global__ void __launch_bounds__(G::GRID_THREADS) grid_prologue_kernel(const __grid_constant__ GArgs a) {
    grid_propose_ahead<G>(a, threadIdx.x);
}

// swap phase + ladder adaptation + output for ladder l by the whole finishing CTA, state staged through smem
template <class G>
__device__ __forceinline__ void grid_ladder_step(const GArgs& a, int l, double* sm, int tid) {
    constexpr int K = G::K, P = G::P, ND = G::ND, THREADS = G::GRID_THREADS;
    LadderWS<K, true> ws;
    ws.carve(sm, reinterpret_cast<int*>(sm + LadderWS<K, true>::DOUBLES));
    for (int j = tid; j < K; j += THREADS) {
        ws.T[j] = a.k.temps[(size_t)l * K + j];
        ws.invT[j] = 1.0 / ws.T[j];
        ws.slot[j] = a.k.slot[(size_t)l * K + j];
    }
    __syncthreads();
    RngKey rng;
    rng.seed = a.k.seed;
    rng.ladder = a.k.ladder0 + (uint32_t)l;
    rng.iter = a.k.rng_iter0 + (uint64_t)a.it;
    ladder_step<G, THREADS, false, true>(a.k, l, a.it, a.k.total_iters0 + a.it + 1, tid, ws, a.k.lpd + (size_t)l * K * ND, rng, a.dbg ? a.dbg + 8 : nullptr);
    if (a.dbg && tid == 0) a.dbg[15] = clock64();
    if (tid < 32) write_traces<G>(a.k, l, a.it, a.k.theta + (size_t)l * K * P, ws.slot, ws.T, ws.lp, tid);
    __syncthreads();
    for (int j = tid; j < K; j += THREADS) {
        a.k.temps[(size_t)l * K + j] = ws.T[j];
        a.k.slot[(size_t)l * K + j] = ws.slot[j];
        a.k.lp[(size_t)l * K + j] = ws.lp[j];
    }
    __syncthreads();
}

template <class G, class S>
__global__ void __launch_bounds__(G::GRID_THREADS, G::GRID_MINB) grid_pass_kernel(const __grid_constant__ GArgs a) {
    constexpr int TB = G::TB, R = G::R, NB = S::NBIG, K = G::K, P = G::P, THREADS = G::GRID_THREADS, NWARP = THREADS / 32, W = G::W;
    static_assert(NB >= 1 && NB <= 2, "grid kernel needs 1 or 2 large declarations");
    extern __shared__ __align__(16) double gsm[];
    double* thT = gsm;                    // [P][TB]
    double* red = gsm + P * TB;           // [NWARP][NB*TB]
    double* s_new = red + NWARP * NB * TB;  // [NB*TB]
    double* s_lad = s_new + NB * TB;      // ladder workspace (LadderWS<K>)
    double* xbuf = s_lad + LadderWS<K, true>::DOUBLES + LadderWS<K, true>::INTS;  // [RING_SLOTS][THREADS] private cp.async rings
    __shared__ int s_last;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nch = a.k.L * K;
    const int q0 = a.pass * TB;
    const long long t_start = a.dbg ? clock64() : 0;
    if (a.dbg && tid == 0 && blockIdx.x == 0) { unsigned long long g_; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(g_)); a.dbg[16 + 2 * ((a.it * a.npass + a.pass) & 31)] = (long long)g_; }
    // One CTA does not stream: it precomputes the increments of the next launch's block proposals, which depend on
    // the rungs' sampler state and on slot-addressed normals only.  The control step then just adds them.
    const bool has_dcta = a.delta != nullptr && a.next_stage >= 0 && G::stage_is_block(a.next_stage) && gridDim.x > 2;
    const int nstream = (int)gridDim.x - (has_dcta ? 1 : 0);
    const bool is_dcta = has_dcta && blockIdx.x == gridDim.x - 1;
    // the next pass may become resident as soon as every CTA of this one has started (it waits for our completion below)
    pdl_launch_dependents();
    // tile mode: the first blocks of this warp's first tile are on their way before anything of the previous pass is needed
    [[maybe_unused]] long wt0 = 0;
    if constexpr (S::TILE) {
        wt0 = (long)warp * nstream + blockIdx.x;
        if (!is_dcta && wt0 < S::big_n(0) / S::TROWS) S::template ring_prime<0>(a.k.D, xbuf, tid, (int)(wt0 * S::TROWS));
    }
    pdl_wait();
    if (is_dcta) { grid_precompute_delta<G>(a, tid); if (a.dbg && tid == 0) a.dbg[7] = clock64() - t_start; }
    // ---- 1. proposed states of this pass's chains -> shared memory, param-major / chain-minor.
    //         CS threads share a row (tile mode: a row group), each owning TBT = TB / CS of the chains.  In tile mode the
    //         chains are dealt to the CS threads with a stride: position (h, c) of the staged parameters holds chain
    //         c * CS + h, so that a thread's c-th chain is one of rungs [c * CS, (c + 1) * CS) -- every thread of a warp holds
    //         one of the CS hottest rungs in the same column of its tile (see bern_logit_tile: the branch that reproduces
    //         the reference's rounding for hot chains is then warp-uniform).  Partial sums travel by position.
    constexpr int CS = S::TILE ? S::CS : G::CSPLIT, TBT = TB / CS, RPC = THREADS / CS;
    auto chain_of_pos = [](int pos) { return S::TILE ? (pos % TBT) * CS + pos / TBT : pos; };
    auto pos_of_chain = [](int q) { return S::TILE ? (q % CS) * TBT + q / CS : q; };
    for (int idx = tid; idx < P * TB; idx += THREADS) {
        const int p = idx / TB, cc = chain_of_pos(idx % TB);
        int q = q0 + cc;
        if (q >= nch) q = nch - 1;
        const int l = q / K, kk = q % K;
        const int sl = a.k.slot[(size_t)l * K + kk];
        thT[idx] = a.k.theta[((size_t)l * K + sl) * P + p];
    }
    __syncthreads();
    // ---- 2. one pass over the data
    const int half = S::TILE ? (lane % CS) : (tid % CS), coff = half * TBT;
    double acc[NB][TBT];
#pragma unroll
    for (int g = 0; g < NB; g++)
#pragma unroll
        for (int cc = 0; cc < TBT; cc++) acc[g][cc] = 0.0;
    if constexpr (S::TILE) {
        // tile mode: warp tiles of TROWS consecutive rows, dealt round-robin over all streaming warps so that the
        // warps with one tile more than the others are spread evenly over the SMs
        constexpr int TR = S::TROWS;
        const long n = S::big_n(0);
        const long ntile = is_dcta ? 0 : n / TR, wstep = (long)nstream * NWARP;
        long wt = wt0;
        int ring_pos = 0;
        for (; wt < ntile; wt += wstep)
            S::template big_eval<0, TBT, S::RT>(a.k.D, thT + coff, (int)(wt * TR), wt + wstep < ntile ? (int)((wt + wstep) * TR) : -1, acc[0], xbuf, ring_pos, tid);
        cp_async_wait<0>();
    } else {
        const long n = S::big_n(0);
        const long ngrp = is_dcta ? 0 : n / R, step = (long)nstream * RPC;
        int ring_pos = 0;
        long grp = (long)blockIdx.x * RPC + tid / CS;
        if (S::RING && grp < ngrp) S::template ring_prime<0>(a.k.D, xbuf, tid, (int)(grp * R));
        for (; grp < ngrp; grp += step)
            S::template big_eval<0, TBT, R>(a.k.D, thT + coff, (int)(grp * R), grp + step < ngrp ? (int)((grp + step) * R) : -1, acc[0], xbuf, ring_pos, tid);
        if (S::RING) cp_async_wait<0>();
        if (!S::RING && blockIdx.x == 0 && tid / CS < (int)(n - ngrp * R)) S::template big_eval<0, TBT, 1>(a.k.D, thT + coff, (int)(ngrp * R + tid / CS), -1, acc[0], xbuf, ring_pos, tid);
    }
    if constexpr (NB == 2) {
        int ring_pos2 = 0;
        const long n = S::big_n(1);
        const long ngrp = is_dcta ? 0 : n / R, step = (long)nstream * RPC;
        for (long grp = (long)blockIdx.x * RPC + tid / CS; grp < ngrp; grp += step)
            S::template big_eval<1, TBT, R>(a.k.D, thT + coff, (int)(grp * R), grp + step < ngrp ? (int)((grp + step) * R) : -1, acc[NB - 1], xbuf, ring_pos2, tid);
        if (blockIdx.x == 0 && tid / CS < (int)(n - ngrp * R)) S::template big_eval<1, TBT, 1>(a.k.D, thT + coff, (int)(ngrp * R + tid / CS), -1, acc[NB - 1], xbuf, ring_pos2, tid);
    }
    const long long t_streamed = a.dbg ? clock64() : 0;
    // ---- 3. CTA reduction (fixed order: lanes of equal `half` by butterfly, warps ascending)
#pragma unroll
    for (int g = 0; g < NB; g++)
#pragma unroll
        for (int cc = 0; cc < TBT; cc++) {
            double v = acc[g][cc];
#pragma unroll
            for (int o = 16; o >= CS; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (lane < CS) red[warp * NB * TB + g * TB + coff + cc] = v;
        }
    __syncthreads();
    double part = 0.0;
    if (tid < NB * TB) {
#pragma unroll
        for (int w = 0; w < NWARP; w++) part += red[w * NB * TB + tid];
    }
    if constexpr (S::TILE) {
        // rows beyond the last whole tile (fewer than TROWS): one row per thread, all TB chains, CTA 0 only
        const long n = S::big_n(0);
        const int rem = (int)(n % S::TROWS);
        if (rem > 0 && blockIdx.x == 0) {
            __syncthreads();
            double ar[TB];
#pragma unroll
            for (int cc = 0; cc < TB; cc++) ar[cc] = 0.0;
            int rp = 0;
            if (tid < rem) S::template big_eval_rem<0, TB, 1>(a.k.D, thT, (int)(n - rem + tid), -1, ar, xbuf, rp, tid);
#pragma unroll
            for (int cc = 0; cc < TB; cc++) {
                double v = ar[cc];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                if (lane == 0) red[warp * NB * TB + cc] = v;
            }
            __syncthreads();
            if (tid < TB) {
#pragma unroll
                for (int w = 0; w < NWARP; w++) part += red[w * NB * TB + tid];
            }
        }
    }
    if (tid < NB * TB) a.partial[(size_t)blockIdx.x * NB * TB + tid] = part;
    // ---- 4. ticket: the last CTA to arrive finishes the stage
    __threadfence();
    __syncthreads();
    if (tid == 0) s_last = (atomicAdd(a.ticket, 1u) == gridDim.x - 1) ? 1 : 0;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    if (a.dbg && tid == 0) { a.dbg[0] = clock64(); a.dbg[5] = t_start; a.dbg[6] = t_streamed; }
    {
        // NB*TB values, each summed over the CTAs' partials in a fixed order by a group of 16 lanes.  All loads of a
        // thread are issued before the first add (one L2 round trip instead of one per batch of four).
        constexpr int GRP = 16, MAXB = 24;  // up to GRP * MAXB = 384 CTAs in registers; beyond that, the loop below
        for (int v = tid / GRP; v < NB * TB; v += THREADS / GRP) {
            const int sub = tid % GRP;
            double x[MAXB];
#pragma unroll
            for (int i = 0; i < MAXB; i++) {
                const int b = sub + i * GRP;
                x[i] = (b < (int)gridDim.x) ? __ldcg(&a.partial[(size_t)b * NB * TB + v]) : 0.0;
            }
            double s = 0.0;
#pragma unroll
            for (int i = 0; i < MAXB; i++) s += x[i];
            for (int b = sub + MAXB * GRP; b < (int)gridDim.x; b += GRP) s += __ldcg(&a.partial[(size_t)b * NB * TB + v]);
#pragma unroll
            for (int o = GRP / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            if (sub == 0) s_new[v] = s;
        }
    }
    // pull the control state of this pass's chains into L1 with one wave of independent prefetches: the control
    // step below is a chain of dependent global loads (slot -> state -> ...), each an L2 round trip otherwise
    {
        const int nchl = min(TB, nch - q0);
        for (int i = tid; i < nchl * 8; i += THREADS) {
            const int q = q0 + i / 8, part = i % 8;
            const int l = q / K, kk = q % K;
            const char* base = nullptr;
            switch (part) {
                case 0: base = (const char*)(a.k.slot + (size_t)l * K + kk); break;
                case 1: base = (const char*)(a.k.lpd + ((size_t)l * K + kk) * G::ND); break;
                case 2: base = (const char*)(a.k.scale + ((size_t)l * K + kk) * G::NU); break;
                case 3: base = (const char*)(a.k.cnt + ((size_t)l * K + kk) * G::NU * 3); break;
                case 4: base = (const char*)(a.pend + (size_t)q * APT_PEND_DOUBLES); break;
                case 5: base = (const char*)(a.k.temps + (size_t)l * K + kk); break;
                case 6: base = (const char*)(a.scratch + (size_t)q * 2 * G::DMAX); break;
                case 7: base = (const char*)(a.scratch + (size_t)q * 2 * G::DMAX + G::DMAX); break;
            }
            prefetch_l1(base);
        }
        for (int i = tid; i < nchl * ((P * 8 + 127) / 128); i += THREADS) {
            const int q = q0 + i / ((P * 8 + 127) / 128), line = i % ((P * 8 + 127) / 128);
            prefetch_l1((const char*)(a.k.theta + (size_t)q * P) + line * 128);
        }
    }
    __syncthreads();
    if (a.dbg && tid == 0) a.dbg[1] = clock64();
    // ---- 5. control step: finish the tempered update of each chain of this pass (one tile of W lanes per chain)
    {
        const int tile = tid / W;
        const int q = q0 + tile;
        if (tile < TB && q < nch) {
            const int l = q / K, kk = q % K;
            Ctx<G> c = grid_ctx<G>(a, l, q, a.it, tid);
            Pend<S::NG> p;
            load_pend<S::NG>(a.pend + (size_t)q * APT_PEND_DOUBLES, p);
            const double* th = c.theta + c.slot[kk] * P;
            double nw[S::NG];
#pragma unroll
            for (int g = 0; g < S::NG; g++) nw[g] = 0.0;
            S::eval(a.k.D, th, 0, c, nw, true, true);  // small groups only (priors, ...)
#pragma unroll
            for (int g = 0; g < S::NG; g++) nw[g] = c.sum(nw[g]);
#pragma unroll
            for (int gi = 0; gi < NB; gi++) nw[S::big_group(gi)] = s_new[gi * TB + pos_of_chain(tile)];
            rw_finish<G, S>(c, a.k.D, kk, 0, p, nw);
        }
    }
    __threadfence_block();
    __syncthreads();
    if (a.dbg && tid == 0) a.dbg[2] = clock64();
    if (a.last_of_iter) {
        for (int l = 0; l < a.k.L; l++) grid_ladder_step<G>(a, l, s_lad, tid);
    }
    if (a.dbg && tid == 0) a.dbg[3] = clock64();
    // ---- 6. propose ahead for the next launch
    grid_propose_ahead<G>(a, tid);
    if (a.dbg) { __syncthreads(); if (tid == 0) { a.dbg[4] = clock64(); unsigned long long g_; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(g_)); a.dbg[17 + 2 * ((a.it * a.npass + a.pass) & 31)] = (long long)g_; } }
    if (tid == 0) *a.ticket = 0u;
}

template <class Kern>
inline cudaError_t launch_pass(Kern kern, int grid, int threads, size_t smem, cudaStream_t st, const GArgs& a) {
    static const bool pdl = getenv("APTB200_NO_PDL") == nullptr;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3((unsigned)threads);
    cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, a);
}

template <class G>
struct GridShape {
    static constexpr size_t smem_bytes(int nbig) {
        return (size_t)(G::P * G::TB + (G::GRID_THREADS / 32) * nbig * G::TB + nbig * G::TB + LadderWS<G::K, true>::DOUBLES + LadderWS<G::K, true>::INTS + (size_t)G::RING_SLOTS * G::GRID_THREADS) * sizeof(double);
    }
};

}  // namespace apt
