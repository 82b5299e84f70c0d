/* apt_module_abi.h — plain-C contract between the front library (libaptb200.so: BUGS parser, lowering, nvcc
 * driver, public C ABI of include/aptb200.h) and a lowered-model module (one nvcc-compiled .so per model that
 * contains the generated device logProb code instantiated into the engine kernels of apt_engine.cuh).
 * No C++ or CUDA types cross this boundary. */
#ifndef APT_MODULE_ABI_H
#define APT_MODULE_ABI_H
#include <stdint.h>

#define APTMOD_ABI_VERSION 7

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
    int n_ladders;        /* independent ladders resident on this device */
    uint32_t ladder0;     /* global id of local ladder 0 (multi-GPU sharding: rank * n_ladders) */
    uint32_t seed;        /* Philox key word 0 */
    int device;           /* CUDA device ordinal */
    const double* temps;  /* n_temps, any order (sorted ascending at build, APT_functions.R:266) */
    int n_temps;
    double ULT;           /* upper limit on temperatures, APT_functions.R:246,497 */
    int monitor_tmax;     /* APT_functions.R:245 */
    int n_arrays;         /* constant/data arrays referenced by the generated code, in generated order */
    const double* const* arrays;
    const int64_t* array_len;
    int n_iarrays;        /* dependency index lists built by the lowering */
    const int* const* iarrays;
    const int64_t* iarray_len;
    const double* theta0; /* initial parameter vector (P) */
    int record;           /* debug: record accept/swap decisions of the next runs */
    int thin_conf, thin2_conf; /* conf$thin, conf$thin2 (APT_functions.R:302-303) */
} aptmod_init;

typedef struct {
    int64_t niter;
    int reset, resetMV, resetTempering, adaptTemps;
    double tuneTemper1, tuneTemper2;
    int thin, thin2;               /* -1 = use conf value (APT_functions.R:351-353) */
    /* test mode: injected randoms (host pointers), see include/aptb200.h "Random-number slots" */
    const double* tabZ;            /* [niter][L][K][NU][DMAX] standard normals */
    const double* tabU;            /* [niter][L][K][NU]       accept uniforms  */
    const double* tabS;            /* [niter][L][K][2]        swap uniforms (rcat, accept) */
    int (*should_abort)(void*);    /* polled between kernel launches (APT_functions.R:415 checkInterrupt) */
    void* abort_arg;
    int iters_per_launch;          /* 0 = default */
    double* samples_out;           /* host destination of this run's mvSamples rows [L][niter/thin][NMON], or null */
    int64_t samples_out_cap;       /* doubles */
    /* Hand-off of this run's mvSamples rows to the caller ON THE DEVICE, piece by piece (the NCCL gather of
     * include/aptb200.h: apt_run_args.gather): after the launch that wrote rows [row0, row1) of this run the module packs
     * them into a contiguous device buffer [L][row1 - row0][NMON] and calls chunk_cb with the engine's stream, on which
     * the packed rows are complete in stream order; whatever the callback enqueues there runs before the next launch.
     * The packed buffer stays valid until the next run.  Non-zero return = error (the run stops with it). */
    int (*chunk_cb)(void* arg, const double* d_packed, int64_t n_doubles, int64_t row0, int64_t row1, void* stream);
    void* chunk_arg;
} aptmod_run_args;

enum {
    APTMOD_SAMPLES = 0, APTMOD_SAMPLES2 = 1, APTMOD_SAMPLES_TMAX = 2, APTMOD_SAMPLES2_TMAX = 3,
    APTMOD_LOGPROBS = 4, APTMOD_TEMPTRAJ = 5, APTMOD_TEMPS = 6, APTMOD_ACCSWAP = 7, APTMOD_LOGPROBTEMPS = 8,
    APTMOD_THETA = 9, APTMOD_SCALE = 10, APTMOD_PROPCOV = 11, APTMOD_DECISIONS = 12, APTMOD_SWAPREC = 13,
    APTMOD_MODEL_THETA = 14
};

enum {
    APTMOD_INFO_P = 0, APTMOD_INFO_K, APTMOD_INFO_NU, APTMOD_INFO_ND, APTMOD_INFO_NMON, APTMOD_INFO_NMON2,
    APTMOD_INFO_DMAX, APTMOD_INFO_ROWS, APTMOD_INFO_ROWS2, APTMOD_INFO_TTROWS, APTMOD_INFO_L,
    APTMOD_INFO_LAUNCHES, APTMOD_INFO_TOTALITERS, APTMOD_INFO_W, APTMOD_INFO_CTA, APTMOD_INFO_ENGINE,
    APTMOD_INFO_BLKSZ,
    APTMOD_INFO_FP64_GFLOPS /* measures: a pure DFMA kernel on the engine's device, timed with CUDA events */
};

/* timing of the last run, measured with CUDA events on the engine's stream */
typedef struct {
    double kernel_ms;     /* sum over launches of the sampling kernels */
    double main_kernel_ms;/* sum over launches of the dominant kernel only */
    int64_t launches;     /* kernels launched by the last run (all kinds) */
    int64_t main_launches;
} aptmod_timing;

#ifdef __cplusplus
}
#endif
#endif
