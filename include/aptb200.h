/*
 * aptb200.h — C ABI of the B200-native adaptive-parallel-tempering engine (libaptb200.so).
 *
 * Drop-in boundary for the hot path of DRJP/nimbleAPT (all reference citations are relative to
 * /root/reference/nimbleAPT/R/APT_functions.R unless another file is named):
 *
 *   reference (R + nimble)                                         this ABI
 *   -------------------------------------------------------------  ---------------------------------------
 *   nimbleCode({...}) / nimbleModel(code, constants, data, inits)   apt_model_create + apt_model_set_array
 *     (vignettes/nimbleAPT-vignette.Rmd:56-72)
 *   conf$addSampler(target, type = "sampler_RW_tempered" |          apt_model_add_sampler
 *     "sampler_RW_block_tempered" | "sampler_slice_tempered" |
 *     "sampler_RW_multinomial_tempered", control = list(...))
 *     (APT:649-660, APT:770-781, APT:898-912, APT:1005-1019; vignette:131-132; demo/APT_demo.R:174-177)
 *   configureMCMC(model, monitors =, thin =) monitors/thin           apt_model_set_monitors / apt_model_set_thin
 *     (APT:302-311)
 *   buildAPT(conf, Temps, monitorTmax, ULT, thinPrintTemps)          apt_build        (lowering + nvcc + device state)
 *     + compileNimble(aptR)   (APT:242-334; vignette:135-137)
 *   cAPT$run(niter, reset, resetMV, resetTempering, simulateAll,     apt_run
 *     time, adaptTemps, printTemps, tuneTemper1, tuneTemper2,
 *     progressBar, thin, thin2)   (APT:336-549)
 *   as.matrix(cAPT$mvSamples), $mvSamples2, $mvSamplesTmax,          apt_get (APT_SAMPLES, ...)
 *     $mvSamples2Tmax, $logProbs, $tempTraj, $accCountSwap, $Temps
 *     (APT:310-317, 386-393, 514-539, 277, 482)
 *   model$calculate()                                                apt_logprob
 *   cAPT$calculateWAIC(nburnin)   (APT:593-635)                      apt_waic
 *
 * Conventions: every entry point returns 0 on success and a non-zero status otherwise; the message is
 * available from apt_last_error() (thread-local).  No C++ exception, longjmp or device pointer crosses this
 * boundary.  All inputs are borrowed for the duration of the call and copied; all outputs are written into
 * caller-allocated host buffers (an R bridge passes REAL(x) of a REALSXP it allocated).  Matrices returned
 * by apt_get are ROW-major [rows][cols] unless stated.  Arrays handed to apt_model_set_array are COLUMN-major
 * (R layout).  One host thread at a time per handle.  There is no CPU fallback: without a B200 apt_build fails.
 *
 * Random-number slots (replaces R's sequential global RNG used at APT:459,463,696,699,843).
 * Philox4x32-10, key = (seed, global ladder id), counter = (iter_lo, iter_hi, unit, rung << 16 | block):
 *   update of sampler unit `unit` at rung `rung` in engine iteration `iter` (a counter that is never reset):
 *     block b < 0x8000   -> two standard normals z[2b], z[2b+1] by Box-Muller on
 *                           u1 = U(w0,w1), u2 = U(w2,w3):  sqrt(-2 ln u1) * (cos, sin)(2 pi u2)
 *     block 0x8000       -> acceptance uniform U(w0,w1) (only consumed when 0 >= logMHR, nimble `decide`)
 *   sampler_slice_tempered (APT:929-964): uniform number j of the update, in the order the reference draws them
 *     (j = 0: rexp(1) = -log U; 1: offset of the initial interval; 2: maxStepsL; 3: first x1; 4, 5, ...: shrinkage
 *     draws), is read from block j / 2 as U(w0,w1) for even j and U(w2,w3) for odd j.  Injected tables are not
 *     consulted by this sampler.
 *   sampler_RW_multinomial_tempered (APT:1054-1098): sub-proposal q (0-based, in the order of the two loops APT:1055-1056)
 *     reads block q: U(w0,w1) is inverted by rbinom (the smallest k with U <= P(X <= k), pmf walked from 0 by its
 *     recurrence; for prob > 0.5 the walk runs on size - X ~ Binom(size, 1 - prob) and size minus its quantile is
 *     returned; size <= 1000; R's rbinom consumes a data-dependent number of sequential draws),
 *     U(w2,w3) is the acceptance uniform.  The recycled u (APT:1040) is pi * U(w0,w1) of block 0x7FFF, drawn at the
 *     sampler's first update.  Injected tables are not consulted by this sampler.
 *   swap proposal ii (0-based) of the iteration: unit = 0xFFFFFFFF, rung field = ii, block 0:
 *     rcat uniform U(w0,w1), acceptance uniform U(w2,w3)
 *   U(hi,lo) = ((hi >> 6) * 2^26 + (lo >> 6) + 0.5) * 2^-52: 52 random bits, strictly inside (0,1).
 * A sampler "unit" is the 0-based position of the sampler in conf$samplerConfs (APT:283 `ss`).
 */
#ifndef APTB200_H
#define APTB200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct apt_model apt_model;   /* model + MCMC configuration under construction */
typedef struct apt_engine apt_engine; /* built engine: compiled module + device-resident state */

/* roles for apt_model_set_array */
enum { APT_ROLE_CONSTANT = 0, APT_ROLE_DATA = 1, APT_ROLE_INIT = 2 };

/* control list of a tempered sampler; field names follow the reference's control keys as the CODE spells
 * them (APT:654-660, APT:775-781).  Initialise with apt_sampler_control_default(). */
typedef struct {
    int log;                    /* control$log            (scalar sampler only)          default FALSE */
    int reflective;             /* control$reflective     (scalar sampler only)          default FALSE */
    int adaptive;               /* control$adaptive                                      default TRUE  */
    int adaptScaleOnly;         /* control$adaptScaleOnly (block sampler only)           default FALSE */
    int adaptInterval;          /* control$adaptInterval                                 default 200   */
    double adaptFactorExponent; /* control$adaptFactorExponent                           default 0.8   */
    double scale;               /* control$scale                                         default 1     */
    int temperPriors;           /* control$temperPriors: -1 = unspecified -> error (APT:660,781) */
    const double* propCov;      /* control$propCov, d x d column-major, NULL = 'identity' (block sampler only) */
    int propCov_dim;            /* d of the matrix passed in propCov (checked against the target, APT:816) */
    double width;               /* control$width          (slice sampler only, APT:911)  default 1   */
    int sliceMaxSteps;          /* control$sliceMaxSteps  (slice sampler only, APT:912)  default 100 */
    int useTempering;           /* control$useTempering   (multinomial sampler only): -1 = unspecified -> error (APT:1015) */
} apt_sampler_control;

typedef struct {
    int monitorTmax;      /* buildAPT(monitorTmax = TRUE)   APT:245 */
    double ULT;           /* buildAPT(ULT = 1e6)            APT:246 */
    int n_ladders;        /* independent ladders resident on this device (the reference runs 1) */
    uint32_t ladder0;     /* global id of the first local ladder (multi-GPU sharding) */
    uint32_t seed;        /* Philox key; replaces set.seed() */
    int device;           /* CUDA device ordinal */
    int record_decisions; /* test hook: keep accept / swap decisions of each run */
    int fuse_links;       /* 1 (default): lower dbinom(prob = ilogit(e), size = 1) to a fused log1p/exp form */
    int lanes_per_chain;  /* 0 = auto; power of two <= 32 */
    int cta_threads;      /* 0 = auto */
    int engine;           /* 0 = auto, 1 = ladder-resident kernel, 2 = grid-wide likelihood passes */
    int rungs_per_pass;   /* grid engine: chains evaluated per pass over the data (0 = auto) */
    int verbose;          /* print the initial ladder like APT:279 and the nvcc command */
    int cluster_ctas;     /* ladder-resident engine: CTAs (SMs) per ladder, 1 | 2 | 4 | 8; 0 = auto (more than one only with
                             few ladders and large sampler families: thread-block cluster, state in global memory) */
    int enable_waic;      /* buildAPT(enableWAIC = TRUE) / configureMCMC(enableWAIC = TRUE)  APT:259,329-333: checks that the
                             model has data nodes and that the monitors are valid for WAIC (mcmc_checkWAICmonitors, APT:1313) */
} apt_build_opts;

/* arguments of cAPT$run (APT:336-348).  Initialise with apt_run_args_default(). */
typedef struct {
    int64_t niter;
    int reset;           /* default TRUE  */
    int resetMV;         /* default FALSE */
    int resetTempering;  /* default FALSE */
    int simulateAll;     /* must be FALSE (simulating from the prior is outside the hot path) */
    int time;            /* TRUE: per-phase CUDA-event times are kept (always on; see apt_get_timing) */
    int adaptTemps;      /* default TRUE  */
    int printTemps;      /* default FALSE: print `iter totalIters-1 rows rows2 Temps` like APT:524 */
    double tuneTemper1;  /* default 10 */
    double tuneTemper2;  /* default 1  */
    int progressBar;     /* default TRUE; honoured by the host wrappers, the library itself prints nothing */
    int thin;            /* default -1 = conf$thin  */
    int thin2;           /* default -1 = conf$thin2 */
    /* extensions */
    int (*should_abort)(void*); /* polled between kernel launches; replaces checkInterrupt() APT:415 */
    void* abort_arg;
    int iters_per_launch;       /* 0 = auto */
    const double* inject_normals;  /* test hook [niter][L][K][NU][DMAX] */
    const double* inject_accept_u; /* test hook [niter][L][K][NU] */
    const double* inject_swap_u;   /* test hook [niter][L][K][2] */
    /* asynchronous copy-back of the thinned output: when non-null, the rows of mvSamples THIS run writes are copied
     * into this host buffer, laid out [n_ladders][niter / thin][n_monitors], in pieces on a second stream while the
     * next piece of the run is still sampling (the run is split into at least 4 launches); complete when apt_run
     * returns.  Page-locked memory makes the copies truly asynchronous.  Rows of earlier runs (reset = FALSE) are not
     * included; apt_get(APT_SAMPLES) still returns everything. */
    double* samples_out;
    int64_t samples_out_capacity;  /* in doubles */
    /* multi-GPU: gather != 0 (needs apt_comm_init; every rank of the communicator calls apt_run with the same niter, thin,
     * gather and gather_root) sends the rows of mvSamples THIS run writes to rank gather_root piece by piece over NCCL
     * while the run goes on; on the root they arrive in samples_out, laid out [ladders of rank 0 | rank 1 | ...][niter / thin]
     * [n_monitors] (ranks may hold different numbers of ladders), complete when apt_run returns.  Other ranks' samples_out
     * is ignored.  The reference's user reads as.matrix(cAPT$mvSamples) in ONE R session (vignette:141): the root is that
     * process.  An interrupted run (should_abort) must be interrupted on every rank at the same piece. */
    int gather;
    int gather_root;
} apt_run_args;

typedef struct {
    double kernel_ms;       /* device time of all sampling kernels of the last run (CUDA events, engine stream) */
    double main_kernel_ms;  /* device time of the dominant kernel only */
    int64_t launches;       /* kernels launched by the last run */
    int64_t main_launches;  /* launches of the dominant kernel */
    double run_wall_ms;     /* host clock around the whole apt_run call */
    double gather_tail_ms;  /* host clock from the end of sampling to the end of apt_run: what was left of the copy-back /
                               gather once the last launch had finished */
} apt_timing;

/* what to fetch with apt_get */
enum {
    APT_SAMPLES = 0,        /* mvSamples      [rows][n_monitors]  rung 1 (T = Temps[1])  APT:515 */
    APT_SAMPLES2 = 1,       /* mvSamples2     [rows2][n_monitors2]                       APT:531 */
    APT_SAMPLES_TMAX = 2,   /* mvSamplesTmax  [rows][n_monitors]  rung nTemps            APT:534 */
    APT_SAMPLES2_TMAX = 3,  /* mvSamples2Tmax                                            APT:537 */
    APT_LOGPROBS = 4,       /* logProbs       [rows]                                     APT:517 */
    APT_TEMPTRAJ = 5,       /* tempTraj       [niter/thin of the last run][nTemps]       APT:516 */
    APT_TEMPS = 6,          /* Temps          [nTemps] current ladder                    APT:496 */
    APT_ACCSWAP = 7,        /* accCountSwap   [nTemps][nTemps] (from, to), last run      APT:482 */
    APT_LOGPROBTEMPS = 8,   /* logProbTemps   [nTemps]                                   APT:447 */
    APT_RUNG_STATES = 9,    /* mvTemps        [nTemps][n_params]                         APT:272 */
    APT_SCALES = 10,        /* sampler scale (slice sampler: width)  [nTemps][n_sampler_units]  APT:743,859,982 */
    APT_BLOCK_STATE = 11,   /* per rung, per block sampler: propCov, chol, scale*chol (d*d each, column-major); slice sampler: sumJumps;
                               multinomial sampler: u, then timesRan, AcceptRates, ScaleShifts, totalAdapted, timesAccepted,
                               ENSwapMatrix, ENSwapDeltaMatrix, RescaleThreshold (d*d each, column-major [iFrom, iTo]) */
    APT_DECISIONS = 12,     /* test hook [niter][nTemps][n_sampler_units] 0/1 (slice sampler: log-probability evaluations of the update mod 256;
                               multinomial sampler: accepted sub-proposals of the update) */
    APT_SWAP_RECORD = 13,   /* test hook [niter][nTemps]  iTo | accepted << 16 */
    APT_MODEL_STATE = 14    /* the model's parameter values after run (rung 1)           APT:548 */
};

/* integer facts about a built engine */
enum {
    APT_INFO_N_PARAMS = 0, APT_INFO_N_TEMPS, APT_INFO_N_SAMPLER_UNITS, APT_INFO_N_LOGPROB_GROUPS,
    APT_INFO_N_MONITORS, APT_INFO_N_MONITORS2, APT_INFO_MAX_BLOCK_DIM, APT_INFO_ROWS, APT_INFO_ROWS2,
    APT_INFO_TEMPTRAJ_ROWS, APT_INFO_N_LADDERS, APT_INFO_LAUNCHES, APT_INFO_TOTAL_ITERS,
    APT_INFO_LANES_PER_CHAIN, APT_INFO_CTA_THREADS, APT_INFO_ENGINE_KIND, APT_INFO_BLOCK_STATE_SIZE,
    APT_INFO_FP64_PEAK_GFLOPS /* not a property but a measurement: FP64 FMA throughput of the engine's device (pure DFMA kernel,
                                 CUDA events), the denominator of the flop roofline of the small-model kernel */
};

const char* apt_last_error(void);
const char* apt_version(void);

/* sizeof of the ABI structs as this library was compiled: 0 apt_sampler_control, 1 apt_build_opts, 2 apt_run_args, 3 apt_timing
 * (a binding that mirrors the structs by hand -- ctypes, an FFI -- checks its own layout against these) */
int64_t apt_abi_sizeof(int which);
void apt_sampler_control_default(apt_sampler_control* c);
void apt_build_opts_default(apt_build_opts* o);
void apt_run_args_default(apt_run_args* a);

/* ---- model + configuration (replaces nimbleCode/nimbleModel/configureMCMC for the supported BUGS subset:
 *      dnorm dunif dgamma dbeta dbinom dbern dpois dexp dcat dmulti dmnorm(cholesky= | prec= | cov= constant), deterministic links,
 *      elements of constant arrays selected by a sampled node, e.g. rfx[k]) */
int apt_model_create(const char* bugs_code, apt_model** out);
void apt_model_free(apt_model* m);
int apt_model_set_array(apt_model* m, const char* name, int role, const double* values, const int64_t* dims, int ndim);
int apt_model_add_sampler(apt_model* m, const char* target, const char* type, const apt_sampler_control* control);
/* A user-supplied deterministic function the model code calls (a nimbleFunction with a `run` function: demo/APT_demo.R:42-71
 * k2theta / k2sigma).  r_source is the R text of the function -- `function(k = double(), v = double(1), ...) { ... }`, or the
 * whole `nimbleFunction(run = function(...) {...})` -- with scalar / vector double arguments, assignments, if / else,
 * return(), arithmetic, comparisons, | & !, ceiling floor round trunc abs exp log sqrt ..., and indexed loads of the vector
 * arguments; it is inlined into the lowered model.  Replaces nimble's compilation of the nimbleFunction. */
int apt_model_add_function(apt_model* m, const char* name, const char* r_source);
int apt_model_remove_samplers(apt_model* m);
int apt_model_set_monitors(apt_model* m, int which /* 1 | 2 */, const char* const* names, int n);
int apt_model_set_thin(apt_model* m, int thin, int thin2);
/* expanded column names of mvSamples (which = 1) / mvSamples2 (which = 2) and of the parameter vector
 * (which = 0), '\n'-separated, written into buf */
int apt_model_names(apt_model* m, int which, char* buf, int64_t buflen);
/* the CUDA source the lowering step emits for this model and ladder size (for inspection / tests) */
int apt_model_lower(apt_model* m, int n_temps, const apt_build_opts* opts, char* buf, int64_t buflen, int64_t* needed);

/* lower + compile the module for this model and ladder size without touching a GPU (nvcc cross-compiles for
 * sm_100a); the path of the cached module is written into path_buf.  apt_build reuses it. */
int apt_model_compile(apt_model* m, int n_temps, const apt_build_opts* opts, char* path_buf, int64_t buflen);

/* ---- buildAPT + compileNimble */
int apt_build(apt_model* m, const double* temps, int n_temps, const apt_build_opts* opts, apt_engine** out);
void apt_free(apt_engine* e);

/* ---- cAPT$run and results */
int apt_run(apt_engine* e, const apt_run_args* args);
int64_t apt_info(apt_engine* e, int what);
/* ladder >= 0: one ladder; ladder = -1: all local ladders, ladder-major.  *n_written receives the count. */
int apt_get(apt_engine* e, int what, int ladder, double* out, int64_t capacity, int64_t* n_written);
/* the same with the layout chosen by the caller: APT_COLUMN_MAJOR (one ladder, ladder >= 0) writes a matrix item the way R
 * stores a matrix, i.e. straight into REAL() of the matrix the R side allocated -- rows x n_monitors for the traces, nTemps x
 * nTemps for accCountSwap, nTemps x n_params for the rung states, ...; traces are transposed on the device.  Vector items
 * (logProbs, Temps, ...) are the same in either layout. */
enum { APT_ROW_MAJOR = 0, APT_COLUMN_MAJOR = 1 };
int apt_get_layout(apt_engine* e, int what, int ladder, int layout, double* out, int64_t capacity, int64_t* n_written);
int apt_get_timing(apt_engine* e, apt_timing* t);
/* whole-model log probability of caller-provided parameter vectors: theta [n][n_params] ->
 * out_total [n] and (optional) out_groups [n][n_logprob_groups] */
int apt_logprob(apt_engine* e, const double* theta, int n, double* out_total, double* out_groups);
/* cAPT$calculateWAIC(nburnin)  (APT:593-635) from the rung-1 samples in mvSamples: nburnin is in iterations before
 * thinning (the rows skipped are ceiling(nburnin / thin of the last run), APT:605).  ladder >= 0: that ladder's samples
 * (the reference has one ladder); -1: the samples of all local ladders pooled.  out[0] = WAIC = -2 (lppd - pWAIC),
 * out[1] = lppd (logAvgProb), out[2] = pWAIC.  Fewer than 2 post-burn-in samples: out[0] = -Inf (APT:607-610).
 * Fails unless the engine was built with enable_waic (the reference prints an error and returns NaN, APT:596-599). */
int apt_waic(apt_engine* e, int64_t nburnin, int ladder, double out[3]);
/* overwrite the model's current parameter values (inits) of one ladder (-1: all) before the next reset run */
int apt_set_inits(apt_engine* e, const double* theta, int ladder);
/* re-upload a constant/data array (same shape) — e.g. new observations for the same model */
int apt_set_array(apt_engine* e, const char* name, const double* values, int64_t n);

/* ---- multi-GPU (one process per GPU; ladders are sharded, no exchange during sampling).
 * NCCL is used only to gather traces and to pool statistics.  unique_id is the 128-byte ncclUniqueId that
 * rank 0 obtained from apt_comm_unique_id and broadcast out of band (e.g. torch.distributed / MPI / file). */
int apt_comm_unique_id(unsigned char id_out[128]);
int apt_comm_init(apt_engine* e, int world_size, int rank, const unsigned char unique_id[128]);
/* all-gather of mvSamples over ranks: out [ladders of rank 0 | rank 1 | ...][rows][n_monitors] on every rank (host memory) */
int apt_comm_allgather_samples(apt_engine* e, int what, double* out, int64_t capacity, int64_t* n_written);
/* gather of mvSamples to ONE rank (grouped ncclSend / ncclRecv over NVLink, one device-to-host copy on the root):
 * out [ladders of rank 0 | rank 1 | ...][rows][n_monitors] on rank `root` only (n_written = 0 and out untouched elsewhere);
 * ranks may hold different numbers of ladders (apt_comm_init exchanges the counts), rows must be equal.  Gathers a whole
 * trace AFTER the run; apt_run_args.gather moves the rows while the run is still sampling.  The reference's
 * user reads as.matrix(cAPT$mvSamples) in one R session (vignette:141), i.e. on one host process. */
int apt_comm_gather_samples(apt_engine* e, int what, int root, double* out, int64_t capacity, int64_t* n_written);
/* sum-all-reduce of a small host vector of pooled statistics (moments, ESS accumulators, swap counts) */
int apt_comm_allreduce_sum(apt_engine* e, double* inout, int64_t n);
int apt_comm_finalize(apt_engine* e);

#ifdef __cplusplus
}
#endif
#endif
