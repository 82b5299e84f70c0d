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

struct GArgs {
    KArgs k;              // k.it0 .. k.it1: the iterations of this launch (a chunk of the run)
    double* partial;      // [grid][2][TB] per-CTA partial sums
    unsigned int* ticket; // arrival counter of the streaming CTAs
    double* pend;         // [L*K][APT_PEND_DOUBLES] pending proposals (old log-probs, Jacobian, old scalar value)
    double* scratch;      // [L*K][2*DMAX]  normals / old block values per chain
    int npass;            // passes per stage: ceil(L * K / TB)
    long long* dbg;       // optional phase timestamps of the control CTA (APTB200_DEBUG_TIMING=1)
    double* delta;        // [L*K][2*DMAX] block-proposal increments precomputed for the next step (+ work space)
    long long* delta_tag; // [L*K][2]  (iteration, timesAdapted) the increments were computed for
    unsigned long long* flags;  // [0] published: sequence number of the latest step whose proposals are in global memory
    unsigned long long seq0;    // sequence number before the first step of this launch
};

// One step = one (iteration, stage, pass): the tempered update of up to TB chains with one pass over the data.
struct Step {
    long long it, next_it;      // iteration (0-based within the run) this step belongs to / the next step's
    int stage, pass;
    int next_stage, next_pass;  // what to propose ahead (-1: nothing, the launch ends)
    int last_of_iter;           // this step completes the iteration: swap phase, ladder adaptation, output
    unsigned long long seq;     // streaming may start when flags[0] >= seq; the control CTA publishes seq + 1
};
template <class G>
__device__ __forceinline__ Step make_step(const GArgs& a, long long n, long long nsteps) {
    const int spi = G::NSTAGE * a.npass;  // steps per iteration
    Step s;
    s.it = a.k.it0 + n / spi;
    const int r = (int)(n % spi);
    s.stage = r / a.npass; s.pass = r % a.npass;
    s.last_of_iter = (r == spi - 1) ? 1 : 0;
    if (n + 1 < nsteps) {
        const int r1 = (int)((n + 1) % spi);
        s.next_it = a.k.it0 + (n + 1) / spi; s.next_stage = r1 / a.npass; s.next_pass = r1 % a.npass;
    } else { s.next_it = 0; s.next_stage = -1; s.next_pass = 0; }
    s.seq = a.seq0 + 1 + (unsigned long long)n;
    return s;
}
template <class S> struct StageTag { using type = S; };

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

// Hand-over between the control CTA and the streaming CTAs of the persistent kernel.  flags[0] ("published") = sequence number
// of the latest step whose proposals are in global memory.  Writer: barrier, one st.release.gpu (cumulative over what the
// barrier ordered before it); readers: one thread polls, fence.acq_rel.gpu (MEMBAR + CCTL.IVALL: the SM's L1 is invalidated),
// barrier.
// The spin itself uses relaxed loads (an acquire load invalidates the SM's L1 every time, which the streaming CTA next door
// pays for: its response and table loads live there) and sleeps between polls; one acquire fence follows the successful poll.
__device__ __forceinline__ unsigned long long ld_relaxed_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned ld_relaxed_u32(const unsigned* p) {
    unsigned v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void fence_acq_rel_gpu() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }
__device__ __forceinline__ void st_release_u64(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
// all threads of the CTA call; returns when *p >= v and everything released before that store is visible to the CTA
#ifndef APT_FLAG_SLEEP_NS
#define APT_FLAG_SLEEP_NS 100
#endif
template <int SLEEP_NS = APT_FLAG_SLEEP_NS>
__device__ __forceinline__ void cta_wait_flag(const unsigned long long* p, unsigned long long v, int tid) {
    if (tid == 0) {
        while (ld_relaxed_u64(p) < v) { if (SLEEP_NS > 0) __nanosleep(SLEEP_NS); }
        fence_acq_rel_gpu();
    }
    __syncthreads();
}
// all threads of the CTA call after their writes (the barrier orders them before thread 0's release, which is cumulative)
__device__ __forceinline__ void cta_publish_flag(unsigned long long* p, unsigned long long v, int tid) {
    __syncthreads();
    if (tid == 0) st_release_u64(p, v);
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

// The control CTA's copy of the control state of one pass in shared memory (a pass that covers whole ladders: TB % K == 0).
// Ladder-shaped like the global arrays (states by slot), so the generic sampler code runs on it through the same Ctx; the
// control step is a chain of dependent loads, each an L2 round trip (~700 cycles) out of global memory and ~30 cycles here.
template <class G>
struct Mirror {
    static constexpr int K = G::K, TB = G::TB, P = G::P, ND = G::ND, NU = G::NU;
    static constexpr bool OK = (TB % K == 0);
    static constexpr int NCH = TB;
    static constexpr int INTS = NCH * (1 + 3 * NU);
    static constexpr int DOUBLES = NCH * (P + ND + 2 + NU + APT_PEND_DOUBLES + 2 * G::DMAX) + (INTS + 1) / 2;
    double *theta, *lpd, *T, *lp, *scale, *pend, *scratch;
    int *slot, *cnt;
    __device__ __forceinline__ void carve(double* d) {
        theta = d; lpd = theta + NCH * P; T = lpd + NCH * ND; lp = T + NCH; scale = lp + NCH; pend = scale + NCH * NU;
        scratch = pend + NCH * APT_PEND_DOUBLES;
        slot = reinterpret_cast<int*>(scratch + NCH * 2 * G::DMAX); cnt = slot + NCH;
    }
    // q0: first chain of the pass (a multiple of K), n: chains of the pass that exist
    __device__ __forceinline__ void load(const GArgs& a, int q0, int n, int tid) {
        constexpr int TH = G::GRID_THREADS;
        for (int i = tid; i < n * P; i += TH) theta[i] = __ldcg(&a.k.theta[(size_t)q0 * P + i]);
        for (int i = tid; i < n * ND; i += TH) lpd[i] = __ldcg(&a.k.lpd[(size_t)q0 * ND + i]);
        for (int i = tid; i < n; i += TH) { T[i] = __ldcg(&a.k.temps[q0 + i]); slot[i] = __ldcg(&a.k.slot[q0 + i]); }
        for (int i = tid; i < n * NU; i += TH) scale[i] = __ldcg(&a.k.scale[(size_t)q0 * NU + i]);
        for (int i = tid; i < n * NU * 3; i += TH) cnt[i] = __ldcg(&a.k.cnt[(size_t)q0 * NU * 3 + i]);
        for (int i = tid; i < n * APT_PEND_DOUBLES; i += TH) pend[i] = __ldcg(&a.pend[(size_t)q0 * APT_PEND_DOUBLES + i]);
        for (int i = tid; i < n * 2 * G::DMAX; i += TH) scratch[i] = __ldcg(&a.scratch[(size_t)q0 * 2 * G::DMAX + i]);
    }
    // what the next pass's streaming CTAs read: the states (slots are stored by the ladder step)
    __device__ __forceinline__ void store_states(const GArgs& a, int q0, int n, int tid) const {
        constexpr int TH = G::GRID_THREADS;
        for (int i = tid; i < n * P; i += TH) a.k.theta[(size_t)q0 * P + i] = theta[i];
        for (int i = tid; i < n; i += TH) a.k.slot[q0 + i] = slot[i];
    }
    // what only the next control step reads
    __device__ __forceinline__ void store_rest(const GArgs& a, int q0, int n, int tid) const {
        constexpr int TH = G::GRID_THREADS;
        for (int i = tid; i < n * ND; i += TH) a.k.lpd[(size_t)q0 * ND + i] = lpd[i];
        for (int i = tid; i < n * NU; i += TH) a.k.scale[(size_t)q0 * NU + i] = scale[i];
        for (int i = tid; i < n * NU * 3; i += TH) a.k.cnt[(size_t)q0 * NU * 3 + i] = cnt[i];
        for (int i = tid; i < n * APT_PEND_DOUBLES; i += TH) a.pend[(size_t)q0 * APT_PEND_DOUBLES + i] = pend[i];
        for (int i = tid; i < n * 2 * G::DMAX; i += TH) a.scratch[(size_t)q0 * 2 * G::DMAX + i] = scratch[i];
    }
    __device__ __forceinline__ void redirect(Ctx<G>& c, int q0, int l, int q) const {
        const int dl = l - q0 / K;
        c.theta = theta + (size_t)dl * K * P;
        c.lpd = lpd + (size_t)dl * K * ND;
        c.slot = slot + dl * K;
        c.T = T + dl * K;
        c.scale = scale + (size_t)dl * K * NU;
        c.cnt = cnt + (size_t)dl * K * NU * 3;
        c.scratch = scratch + (size_t)(q - q0) * 2 * G::DMAX;
    }
};

// Increments of the next step's block proposals, private to the control CTA (shared memory): [TB][2 * DMAX] doubles (increments,
// work space) and the (iteration, timesAdapted) pair they were computed for.
struct DeltaBuf {
    double* inc;
    long long* tag;
};

// proposals of the next step for control tile tid / W; with a mirror (and the same chains next) they are written into it
template <class G>
__device__ __forceinline__ void grid_propose_ahead(const GArgs& a, const Step& s, int tid, const Mirror<G>* m = nullptr, const DeltaBuf* db = nullptr) {
    constexpr int K = G::K, W = G::W, TB = G::TB;
    if (s.next_stage < 0) return;
    const int tile = tid / W;
    if (tile >= TB) return;
    const int q = s.next_pass * TB + tile;
    if (q >= a.k.L * K) return;
    const int l = q / K, kk = q % K;
    Ctx<G> c = grid_ctx<G>(a, l, q, s.next_it, tid);
    double* pend = a.pend + (size_t)q * APT_PEND_DOUBLES;
    if (m) { m->redirect(c, s.next_pass * TB, l, q); pend = m->pend + (size_t)tile * APT_PEND_DOUBLES; }
    const double* delta = nullptr;
    if (db && G::stage_is_block(s.next_stage)) {
        // valid only if computed for this very step and if the rung has not adapted since (APT:859-868)
        const long long t_it = db->tag[tile * 2], t_ad = db->tag[tile * 2 + 1];
        const int nad = c.cnt[(kk * G::NU + G::stage_unit0(s.next_stage)) * 3 + 2];
        if (t_it == s.next_it * 4096 + s.next_stage * 64 + s.next_pass && t_ad == (long long)nad) delta = db->inc + (size_t)tile * 2 * G::DMAX;
    }
    G::propose_stage(s.next_stage, c, a.k.D, kk, pend, delta);
}

// the control CTA: increments of the NEXT step's block proposals, computed while the others stream
template <class G>
__device__ __forceinline__ void grid_precompute_delta(const GArgs& a, const Step& s, int tid, const DeltaBuf& db) {
    constexpr int K = G::K, W = G::W, TB = G::TB;
    const int tile = tid / W;
    if (tile >= TB) return;
    const int q = s.next_pass * TB + tile;
    if (q >= a.k.L * K) return;
    const int l = q / K, kk = q % K;
    Ctx<G> c = grid_ctx<G>(a, l, q, s.next_it, tid);
    double* out = db.inc + (size_t)tile * 2 * G::DMAX;
    const int nad = c.cnt[(kk * G::NU + G::stage_unit0(s.next_stage)) * 3 + 2];
    G::delta_stage(s.next_stage, c, kk, out + G::DMAX, out);
    if (c.rank == 0) {
        db.tag[tile * 2] = s.next_it * 4096 + s.next_stage * 64 + s.next_pass;
        db.tag[tile * 2 + 1] = (long long)nad;
    }
}

// swap phase (+ ladder adaptation + output) for ladder l by the whole control CTA.  PHASE bits as in ladder_step: 8 = what
// depends on the iteration and on the ladder before the step only (ladder, 1 / T, slots, swap uniforms and their logs: done
// while the data streams), 1 = pSwapCalc, the K swap proposals, the output rows that read the states, slots stored (4: bit 8
// has been done), 2 = ladder adaptation, tempTraj, ladder stored.
template <class G, int PHASE>
__device__ __forceinline__ void grid_ladder_step(const GArgs& a, const Step& s, int l, double* sm, int tid, const Mirror<G>* m, int q0) {
    constexpr int K = G::K, P = G::P, ND = G::ND, THREADS = G::GRID_THREADS;
    LadderWS<K, true> ws;
    ws.carve(sm, reinterpret_cast<int*>(sm + LadderWS<K, true>::DOUBLES));
    const int dl = m ? l - q0 / K : 0;
    const double* lpd_l = m ? m->lpd + (size_t)dl * K * ND : a.k.lpd + (size_t)l * K * ND;
    const double* theta_l = m ? m->theta + (size_t)dl * K * P : a.k.theta + (size_t)l * K * P;
    RngKey rng;
    rng.seed = a.k.seed;
    rng.ladder = a.k.ladder0 + (uint32_t)l;
    rng.iter = a.k.rng_iter0 + (uint64_t)s.it;
    if constexpr ((PHASE & 8) != 0 || ((PHASE & 1) != 0 && (PHASE & 4) == 0)) {
        for (int j = tid; j < K; j += THREADS) {
            ws.T[j] = m ? m->T[dl * K + j] : a.k.temps[(size_t)l * K + j];
            ws.invT[j] = 1.0 / ws.T[j];
            ws.slot[j] = m ? m->slot[dl * K + j] : a.k.slot[(size_t)l * K + j];
        }
        __syncthreads();
    }
    if constexpr ((PHASE & 8) != 0) ladder_step<G, THREADS, false, true, 8>(a.k, l, s.it, a.k.total_iters0 + s.it + 1, tid, ws, lpd_l, rng, nullptr);
    if constexpr ((PHASE & 1) != 0) {
        ladder_step<G, THREADS, false, true, (PHASE & 5)>(a.k, l, s.it, a.k.total_iters0 + s.it + 1, tid, ws, lpd_l, rng, a.dbg ? a.dbg + 8 : nullptr);
        // the rows that read the states are written now: the next proposals overwrite the states
        if (tid < 32) write_traces<G, 1>(a.k, l, s.it, theta_l, ws.slot, ws.T, ws.lp, tid);
        __syncthreads();
        for (int j = tid; j < K; j += THREADS) {
            if (m) m->slot[dl * K + j] = ws.slot[j];
            a.k.slot[(size_t)l * K + j] = ws.slot[j];
        }
    }
    if constexpr ((PHASE & 2) != 0) {
        ladder_step<G, THREADS, false, true, 2>(a.k, l, s.it, a.k.total_iters0 + s.it + 1, tid, ws, lpd_l, rng, nullptr);
        if (tid < 32) write_traces<G, 2>(a.k, l, s.it, theta_l, ws.slot, ws.T, ws.lp, tid);
        __syncthreads();
        for (int j = tid; j < K; j += THREADS) {
            if (m) m->T[dl * K + j] = ws.T[j];
            a.k.temps[(size_t)l * K + j] = ws.T[j];
            a.k.lp[(size_t)l * K + j] = ws.lp[j];
        }
    }
    __syncthreads();
}

// shared-memory layout of the persistent kernel (both roles)
template <class G>
struct GridSmem {
    static constexpr int K = G::K, P = G::P, TB = G::TB, NWARP = G::GRID_THREADS / 32, NBMAX = 2;
    static constexpr int O_RED = P * TB, O_NEW = O_RED + NWARP * NBMAX * TB, O_LAD = O_NEW + NBMAX * TB;
    static constexpr int O_FRAG = (O_LAD + LadderWS<K, true>::DOUBLES + LadderWS<K, true>::INTS + 1) & ~1;  // 16-byte aligned
    static constexpr int O_XBUF = (O_FRAG + G::FRAG_DOUBLES + 1) & ~1;
    static constexpr size_t MIRROR = (size_t)((Mirror<G>::OK ? Mirror<G>::DOUBLES + 2 : 0) + 1) & ~(size_t)1;  // control CTA: mirror, then increments
    static constexpr size_t DELTA = (size_t)TB * 2 * G::DMAX + (size_t)TB * 2;
    static constexpr size_t XBUF = (size_t)G::RING_SLOTS * G::GRID_THREADS > MIRROR + DELTA ? (size_t)G::RING_SLOTS * G::GRID_THREADS : MIRROR + DELTA;
    static constexpr size_t BYTES = ((size_t)O_XBUF + XBUF) * sizeof(double);
};

// ---- a streaming CTA's part of one step: stage the proposed states, one pass over the data, partial sums, ticket
template <class G, class S>
__device__ __forceinline__ void grid_stream_step(const GArgs& a, const Step& s, double* gsm, int nstream) {
    constexpr int TB = G::TB, R = G::R, NB = S::NBIG, K = G::K, P = G::P, THREADS = G::GRID_THREADS, NWARP = THREADS / 32;
    static_assert(NB >= 1 && NB <= 2, "grid kernel needs 1 or 2 large declarations");
    using SM = GridSmem<G>;
    double* thT = gsm;                    // [P][TB]
    double* red = gsm + SM::O_RED;        // [NWARP][NB*TB]
    double* thF = gsm + SM::O_FRAG;       // [FRAG_DOUBLES] parameters in MMA fragment order
    double* xbuf = gsm + SM::O_XBUF;      // [RING_SLOTS][THREADS] private cp.async rings
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nch = a.k.L * K;
    const int q0 = s.pass * TB;
    constexpr int CS = S::TILE ? S::CS : G::CSPLIT, TBT = TB / CS, RPC = THREADS / CS;
    auto chain_of_pos = [](int pos) { return S::TILE ? (pos % TBT) * CS + pos / TBT : pos; };
    // tile mode: the first blocks of this warp's first tile are on their way before anything of the previous step is needed
    [[maybe_unused]] long wt0 = 0;
    if constexpr (S::TILE) {
        wt0 = (long)warp * nstream + blockIdx.x;
        if (wt0 < S::big_n(0) / S::TROWS) S::template ring_prime<0>(a.k.D, xbuf, tid, (int)(wt0 * S::TROWS));
    }
    // the proposals of this step are published by the control CTA
    cta_wait_flag(a.flags, s.seq, tid);
    // ---- 1. proposed states of this pass's chains -> shared memory, param-major / chain-minor.
    //         CS threads share a row (tile mode: a row group), each owning TBT = TB / CS of the chains.  In tile mode the
    //         chains are dealt to the CS threads with a stride: position (h, c) of the staged parameters holds chain
    //         c * CS + h, so that a thread's c-th chain is one of rungs [c * CS, (c + 1) * CS) -- every thread of a warp holds
    //         one of the CS hottest rungs in the same column of its tile (see bern_logit_tile: the branch that reproduces
    //         the reference's rounding for hot chains is then warp-uniform).  Partial sums travel by position.
    for (int idx = tid; idx < P * TB; idx += THREADS) {
        const int p = idx / TB, cc = chain_of_pos(idx % TB);
        int q = q0 + cc;
        if (q >= nch) q = nch - 1;
        const int l = q / K, kk = q % K;
        const int sl = __ldcg(&a.k.slot[(size_t)l * K + kk]);
        thT[idx] = __ldcg(&a.k.theta[((size_t)l * K + sl) * P + p]);
    }
    __syncthreads();
    if constexpr (S::MMA) {
        S::mma_stage(thT, thF, tid);
        __syncthreads();
    }
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
        const long ntile = n / TR, wstep = (long)nstream * NWARP;
        long wt = wt0;
        int ring_pos = 0;
        for (; wt < ntile; wt += wstep)
            S::template big_eval<0, TBT, S::RT>(a.k.D, thT + coff, (int)(wt * TR), wt + wstep < ntile ? (int)((wt + wstep) * TR) : -1, acc[0], xbuf, ring_pos, tid);
        cp_async_wait<0>();
    } else {
        const long n = S::big_n(0);
        const long ngrp = n / R, step = (long)nstream * RPC;
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
        const long ngrp = n / R, step = (long)nstream * RPC;
        for (long grp = (long)blockIdx.x * RPC + tid / CS; grp < ngrp; grp += step)
            S::template big_eval<1, TBT, R>(a.k.D, thT + coff, (int)(grp * R), grp + step < ngrp ? (int)((grp + step) * R) : -1, acc[NB - 1], xbuf, ring_pos2, tid);
        if (blockIdx.x == 0 && tid / CS < (int)(n - ngrp * R)) S::template big_eval<1, TBT, 1>(a.k.D, thT + coff, (int)(ngrp * R + tid / CS), -1, acc[NB - 1], xbuf, ring_pos2, tid);
    }
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
    if (tid < NB * TB) a.partial[(size_t)tid * gridDim.x + blockIdx.x] = part;  // value-major: the control CTA reads a value's partials coalesced
    // ---- 4. arrival: partial sums released, ticket taken
    __threadfence();
    __syncthreads();
    if (tid == 0) atomicAdd(a.ticket, 1u);
}

// ---- the control CTA's part of one step
template <class G, class S>
__device__ __forceinline__ void grid_control_step(const GArgs& a, const Step& s, double* gsm, int nstream) {
    constexpr int TB = G::TB, NB = S::NBIG, K = G::K, P = G::P, THREADS = G::GRID_THREADS, W = G::W;
    using SM = GridSmem<G>;
    double* s_new = gsm + SM::O_NEW;      // [NB*TB]
    double* s_lad = gsm + SM::O_LAD;      // ladder workspace (LadderWS<K>)
    double* xbuf = gsm + SM::O_XBUF;      // the mirror, then the increments
    const int tid = threadIdx.x;
    const int nch = a.k.L * K;
    const int q0 = s.pass * TB;
    constexpr int CS = S::TILE ? S::CS : G::CSPLIT, TBT = TB / CS;
    auto pos_of_chain = [](int q) { return S::TILE ? (q % CS) * TBT + q / CS : q; };
    const long long t_start = a.dbg ? clock64() : 0;
    // =============== while the others stream: everything that does not depend on the sums they are computing
    // (control steps are chains of dependent FP64 operations at ~36 cycles each on one SM: what can leave the critical path, does)
    // -- the increments of the next step's block proposals (rung's sampler state and slot-addressed normals only)
    DeltaBuf db;
    db.inc = xbuf + SM::MIRROR;
    db.tag = reinterpret_cast<long long*>(db.inc + (size_t)TB * 2 * G::DMAX);
    const bool want_delta = a.delta != nullptr && s.next_stage >= 0 && G::stage_is_block(s.next_stage);
    if (want_delta) grid_precompute_delta<G>(a, s, tid, db);
    const long long t_delta = a.dbg ? clock64() : 0;
    // -- the control state of this pass into shared memory
    constexpr bool MIR = Mirror<G>::OK;
    Mirror<G> mir;
    const int nchl = min(TB, nch - q0);
    if constexpr (MIR) {
        mir.carve(xbuf);
        mir.load(a, q0, nchl, tid);
    }
    const Mirror<G>* mp = MIR ? &mir : nullptr;
    __syncthreads();
    // -- the small dependent groups of the proposals (priors, ...), the uniform of the decision
    const int tile = tid / W;
    const int q = q0 + tile;
    const bool has_chain = tile < TB && q < nch;
    const int l_ = has_chain ? q / K : 0, kk = has_chain ? q % K : 0;
    Ctx<G> c = grid_ctx<G>(a, l_, has_chain ? q : 0, s.it, tid);
    const double* pend = a.pend + (size_t)(has_chain ? q : 0) * APT_PEND_DOUBLES;
    if constexpr (MIR) { mir.redirect(c, q0, l_, has_chain ? q : q0); pend = mir.pend + (size_t)(has_chain ? tile : 0) * APT_PEND_DOUBLES; }
    Pend<S::NG> p;
    double nw[S::NG];
    double u_acc = -1.0;
#pragma unroll
    for (int g = 0; g < S::NG; g++) nw[g] = 0.0;
    if (has_chain) {
        load_pend<S::NG>(pend, p);
        const double* th = c.theta + c.slot[kk] * P;
        S::eval(a.k.D, th, 0, c, nw, true, true);  // small groups only
#pragma unroll
        for (int g = 0; g < S::NG; g++) nw[g] = c.sum(nw[g]);
        u_acc = rw_accept_uniform<G, S>(c, kk, 0);
    }
    // -- the ladder, 1 / T and the swap uniforms of the iteration (one ladder: the workspace stays put until the swaps)
    const bool defer = s.last_of_iter && a.k.L == 1;
    if (defer) grid_ladder_step<G, 8>(a, s, 0, s_lad, tid, mp, q0);
    __syncthreads();
    const long long t_loaded = a.dbg ? clock64() : 0;
    // =============== the critical path: from the last streaming CTA's arrival to the published proposals
    if (tid == 0) {
        while (ld_relaxed_u32(a.ticket) < (unsigned)nstream) {}
        fence_acq_rel_gpu();
        *a.ticket = 0u;  // the streaming CTAs take the next step's tickets only after this step has published
    }
    __syncthreads();
    if (a.dbg && tid == 0) { a.dbg[0] = clock64(); a.dbg[5] = t_start; a.dbg[6] = t_loaded; a.dbg[7] = t_delta - t_start; }
    {
        // NB*TB values, each summed over the CTAs' partials in a fixed order by a group of 16 lanes.  All loads of a
        // thread are issued before the first add (one L2 round trip instead of one per batch of four).
        constexpr int GRP = 16, MAXB = 24;  // up to GRP * MAXB = 384 CTAs in registers; beyond that, the loop below
        for (int v = tid / GRP; v < NB * TB; v += THREADS / GRP) {
            const int sub = tid % GRP;
            const double* pv = a.partial + (size_t)v * gridDim.x;
            double x[MAXB];
#pragma unroll
            for (int i = 0; i < MAXB; i++) {
                const int b = sub + i * GRP;
                x[i] = (b < nstream) ? __ldcg(&pv[b]) : 0.0;
            }
            double sum = 0.0;
#pragma unroll
            for (int i = 0; i < MAXB; i++) sum += x[i];
            for (int b = sub + MAXB * GRP; b < nstream; b += GRP) sum += __ldcg(&pv[b]);
#pragma unroll
            for (int o = GRP / 2; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
            if (sub == 0) s_new[v] = sum;
        }
    }
    __syncthreads();
    if (a.dbg && tid == 0) a.dbg[1] = clock64();
    // ---- finish the tempered update of each chain of this pass (one tile of W lanes per chain)
    if (has_chain) {
#pragma unroll
        for (int gi = 0; gi < NB; gi++) nw[S::big_group(gi)] = s_new[gi * TB + pos_of_chain(tile)];
        rw_finish<G, S>(c, a.k.D, kk, 0, p, nw, u_acc);
    }
    __syncthreads();
    if (a.dbg && tid == 0) a.dbg[2] = clock64();
    // ---- swap phase (after the last stage of an iteration).  With one ladder the ladder adaptation and tempTraj wait
    //      until the next proposals are published: neither they nor the streaming CTAs depend on them.
    if (s.last_of_iter) {
        if (defer) grid_ladder_step<G, 5>(a, s, 0, s_lad, tid, mp, q0);
        else
            for (int l = 0; l < a.k.L; l++) {
                const bool mine = MIR && l >= q0 / K && l < (q0 + nchl) / K;
                grid_ladder_step<G, 3>(a, s, l, s_lad, tid, mine ? mp : nullptr, q0);
            }
    }
    if (a.dbg && tid == 0) a.dbg[3] = clock64();
    // ---- propose ahead for the next step (into the mirror when it works on the same chains), publish
    const bool same_next = MIR && s.next_stage >= 0 && s.next_pass == s.pass;
    if constexpr (MIR) {
        if (!same_next) { mir.store_states(a, q0, nchl, tid); mir.store_rest(a, q0, nchl, tid); __syncthreads(); }
    }
    grid_propose_ahead<G>(a, s, tid, same_next ? mp : nullptr, want_delta ? &db : nullptr);
    if constexpr (MIR) {
        if (same_next) { __syncthreads(); mir.store_states(a, q0, nchl, tid); }
    }
    cta_publish_flag(a.flags, s.seq + 1, tid);
    if (a.dbg && tid == 0) a.dbg[4] = clock64();
    // =============== what nothing on the critical path waits for
    if constexpr (MIR) {
        if (same_next) mir.store_rest(a, q0, nchl, tid);
    }
    if (defer) grid_ladder_step<G, 2>(a, s, 0, s_lad, tid, mp, q0);
    __syncthreads();
    if (a.dbg && tid == 0) a.dbg[15] = clock64();
}

// The persistent kernel of the grid engine: one cooperative launch (all CTAs resident, 2 per SM) runs a whole chunk of
// iterations.  The last CTA is the control CTA, the others stream; they meet through the ticket (streaming -> control) and
// the published sequence number (control -> streaming), never through a kernel boundary: no launch gap, and the control
// path -- ~50 KB of code executed once per step -- stays in one SM's instruction cache.
template <class G>
__global__ void __launch_bounds__(G::GRID_THREADS, G::GRID_MINB) grid_run_kernel(const __grid_constant__ GArgs a) {
    extern __shared__ __align__(16) double gsm[];
    const int nstream = (int)gridDim.x - 1;
    const bool is_cc = blockIdx.x == gridDim.x - 1;
    const long long nsteps = (a.k.it1 - a.k.it0) * G::NSTAGE * a.npass;
    if (is_cc) {
        // the proposals of the first step
        Step s0;
        s0.it = 0; s0.stage = -1; s0.pass = -1; s0.last_of_iter = 0; s0.seq = a.seq0;
        s0.next_it = a.k.it0; s0.next_stage = 0; s0.next_pass = 0;
        grid_propose_ahead<G>(a, s0, threadIdx.x, nullptr, nullptr);
        cta_publish_flag(a.flags, a.seq0 + 1, threadIdx.x);
    }
    for (long long n = 0; n < nsteps; n++) {
        const Step s = make_step<G>(a, n, nsteps);
        G::with_stage(s.stage, [&](auto tag) {
            using S = typename decltype(tag)::type;
            if (is_cc) grid_control_step<G, S>(a, s, gsm, nstream);
            else grid_stream_step<G, S>(a, s, gsm, nstream);
        });
    }
}

template <class G>
struct GridShape {
    static constexpr size_t smem_bytes(int) { return GridSmem<G>::BYTES; }
};

}  // namespace apt
