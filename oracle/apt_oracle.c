/*
 * oracle/apt_oracle.c — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this.
 *
 * Single-threaded-per-ladder CPU restatement of the reference's adaptive-parallel-tempering hot path:
 *   iteration skeleton           /root/reference/nimbleAPT/R/APT_functions.R:413-546
 *   run prologue (reset/continue) APT_functions.R:350-393
 *   pSwapCalc                    APT_functions.R:556-576
 *   sampler_RW_tempered          APT_functions.R:649-761
 *   sampler_RW_block_tempered    APT_functions.R:770-889
 *   sampler_slice_tempered       APT_functions.R:898-995
 *   sampler_RW_multinomial_tempered APT_functions.R:1005-1145
 *   calculateWAIC                APT_functions.R:593-635
 *   sampler indexing             APT_functions.R:281-287
 * plus the nimble behaviours those lines call into (calculateDiff, getLogProb, decide, decideAndJump,
 * setAndCalculateDiff, calcAdaptationFactor, rmnorm_chol, rcat, nimCopy, initializeModel), restated from
 * nimble's documented semantics [NIMBLE-UNVERIFIED: nimble is an un-vendored, un-pinned dependency,
 * nimbleAPT/DESCRIPTION:19-21; SURVEY.md §2b].
 *
 * PARITY UNPINNED: the reference ships no tests, golden vectors or fixtures and cannot be
 * run here (no R, no nimble).  This restatement is pinned instead by scipy (densities), analytic posteriors
 * and invariants taken from the reference code (tests/test_oracle_*.py).
 *
 * The model is an interpreted node program ("IR") produced by oracle/bugs_front.py — a front end that is
 * independent of the product's C++ parser/lowering — and is evaluated NODE BY NODE with stored per-node
 * logProbs, the way nimble evaluates a model.  Randoms are slot-addressed (philox.h) or read from a table.
 *
 * Differences from the reference kept on purpose and result-neutral: every rung keeps its own model/mvSaved
 * buffers, so the whole-model nimCopy's of APT:438-439,446,466-468 become pointer swaps (this makes the CPU
 * baseline FASTER than compiled nimble would be, i.e. favourable to the reference).
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <unistd.h>
#include "philox.h"
#include "rmath_port.h"

#define DREC 80 /* ints per declaration record */
#define MAXARGS 6
#define ARGREC 8

enum { OP_END = 0, OP_CONST, OP_LOOP, OP_LOAD0, OP_LOAD1, OP_LOAD2, OP_ADD, OP_SUB, OP_MUL, OP_DIV, OP_POW, OP_NEG,
       OP_EXP, OP_LOG, OP_ABS, OP_SQRT, OP_ILOGIT, OP_LOGIT, OP_PHI, OP_ICLOGLOG, OP_CLOGLOG, OP_LOG1P, OP_STEP,
       OP_MIN, OP_MAX, OP_INPROD, OP_SUM, OP_LGAMMA,
       /* bodies of user-supplied deterministic functions (demo/APT_demo.R:42-71) */
       OP_CEIL, OP_FLOOR, OP_ROUND, OP_TRUNC, OP_LT, OP_GT, OP_LE, OP_GE, OP_EQ, OP_NE, OP_OR, OP_AND, OP_NOT, OP_SEL };
enum { D_NORM = 0, D_UNIF, D_GAMMA, D_BETA, D_BINOM, D_POIS, D_MNORM_CHOL, D_EXP, D_CAT, D_MULTI };

/* declaration record layout (ints) */
enum { R_KIND = 0, R_DIST, R_NLOOP, R_LO0, R_LO1, R_LO2, R_HI0, R_HI1, R_HI2, R_VAR, R_LNDIM, R_LIDX0, R_LIDX1,
       R_LRDIM, R_LRLO, R_LRHI, R_NARGS, R_ARGS /* MAXARGS*ARGREC */ };
/* arg record: [kind, a,b,c,d,e,f,g]
   kind 0: scalar expr  a=code offset
   kind 1: vector slice a=var b=ndim c=rangedim d=lo e=hi f=code offset of the other dim's index (-1 none)
   kind 2: matrix slice a=var b=lo1 c=hi1 d=lo2 e=hi2 */

typedef struct {
    int nvars, nmut, ntot, ndecl, ncode, nconst, nlp;
    int *var_off, *var_nrow;
    int* decl;
    int* code;
    double* consts;
    int* lp_off;
    double* arena0; /* initial values of the whole arena (const/data part is read from here) */
} orc_model;

typedef struct {
    int kind, d;
    int* tgt;
    int ncalc;
    int *calc_decl, *calc_inst;
    int ntnode;
    int *tn_decl, *tn_inst;
    int log_scale, reflective, adaptive, adapt_scale_only, adapt_interval, temper_priors;
    double adapt_exp, scale0;
    double* propcov0;
    int lower_code, upper_code;
    double ntotal;           /* sampler_RW_multinomial_tempered: Ntotal <- sum(values(model, target)) at setup (APT:1024) */
    int max_steps, discrete; /* sampler_slice_tempered: control$sliceMaxSteps (APT:912), model$isDiscrete(target) (APT:926) */
    int ncopyv, ncopylp;
    int *copyv, *copylp;
} orc_sspec;

typedef struct {
    double scale, gamma1; /* slice sampler: scale = width */
    double sumJumps;      /* slice sampler (APT:925) */
    int timesRan, timesAccepted, timesAdapted;
    double *propCov, *chol, *cholScaled, *empir;
    double u;   /* multinomial sampler: recycled uniform on (0, pi) (APT:1040); < 0 = not drawn yet */
    double* mn; /* multinomial sampler: 8 matrices d x d (APT:1029-1036) */
} orc_sstate;

typedef struct {
    double *mut, *lp;
} orc_buf;

typedef struct {
    const orc_model* m;
    int K, NS;
    const orc_sspec* specs;
    uint32_t seed, ladder;
    double ULT;
    int monitorTmax;
    int nmon, nmon2;
    int *mon, *mon2; /* arena idx (>=0) or -(lp idx)-1 */
    double *Temps, *invTemps, *TempsCurrent, *logProbTemps, *pSwap, *temporary;
    double* accCountSwap;
    orc_buf* rung;   /* mvTemps rows */
    orc_buf* saved;  /* per-rung mvSaved */
    orc_buf model;   /* the "model" object between runs */
    orc_sstate* ss;  /* K*NS, index ss + tt*NS  (APT:283) */
    long totalIters;
    uint64_t rngIter;
    int resetAnyway;
    /* outputs */
    long nrows, nrows2, cap, cap2, ttrows;
    double *samples, *samples2, *samplesTmax, *samples2Tmax, *logProbs, *tempTraj;
    /* random table / decision record (optional, per run) */
    const double *tabZ, *tabU, *tabS;
    int tabDmax;
    unsigned char* recDec;
    int* recSwap;
    long itRun;
    double work[64];
} orc_ladder;

/* ------------------------------------------------------------------ interpreter */
typedef struct {
    const orc_model* m;
    double* mut;
    double* lp;
    double L[3];
} orc_ctx;

static inline double* vref(orc_ctx* c, int idx) { return idx < c->m->nmut ? &c->mut[idx] : &c->m->arena0[idx]; }

static double slice_get(orc_ctx* c, const int* a, double other, int e) {
    /* element e (0-based) of a vector slice described by arg record a (kind 1) */
    const orc_model* m = c->m;
    int var = a[1], ndim = a[2], rdim = a[3], lo = a[4];
    int off = m->var_off[var];
    if (ndim == 1) return *vref(c, off + (lo - 1 + e));
    int nrow = m->var_nrow[var];
    if (rdim == 0) return *vref(c, off + (lo - 1 + e) + ((int)other - 1) * nrow);
    return *vref(c, off + ((int)other - 1) + (lo - 1 + e) * nrow);
}

static double eval(orc_ctx* c, int pc) {
    const orc_model* m = c->m;
    const int* code = m->code;
    double st[64];
    int sp = 0;
    for (;;) {
        int op = code[pc], a = code[pc + 1];
        pc += 2;
        switch (op) {
            case OP_END: return st[sp - 1];
            case OP_CONST: st[sp++] = m->consts[a]; break;
            case OP_LOOP: st[sp++] = c->L[a]; break;
            case OP_LOAD0: st[sp++] = *vref(c, m->var_off[a]); break;
            case OP_LOAD1: st[sp - 1] = *vref(c, m->var_off[a] + (int)st[sp - 1] - 1); break;
            case OP_LOAD2:
                st[sp - 2] = *vref(c, m->var_off[a] + ((int)st[sp - 2] - 1) + ((int)st[sp - 1] - 1) * m->var_nrow[a]);
                sp--;
                break;
            case OP_ADD: st[sp - 2] += st[sp - 1]; sp--; break;
            case OP_SUB: st[sp - 2] -= st[sp - 1]; sp--; break;
            case OP_MUL: st[sp - 2] *= st[sp - 1]; sp--; break;
            case OP_DIV: st[sp - 2] /= st[sp - 1]; sp--; break;
            case OP_POW: st[sp - 2] = pow(st[sp - 2], st[sp - 1]); sp--; break;
            case OP_MIN: st[sp - 2] = fmin(st[sp - 2], st[sp - 1]); sp--; break;
            case OP_MAX: st[sp - 2] = fmax(st[sp - 2], st[sp - 1]); sp--; break;
            case OP_NEG: st[sp - 1] = -st[sp - 1]; break;
            case OP_EXP: st[sp - 1] = exp(st[sp - 1]); break;
            case OP_LOG: st[sp - 1] = log(st[sp - 1]); break;
            case OP_ABS: st[sp - 1] = fabs(st[sp - 1]); break;
            case OP_SQRT: st[sp - 1] = sqrt(st[sp - 1]); break;
            case OP_ILOGIT: st[sp - 1] = 1.0 / (1.0 + exp(-st[sp - 1])); break;
            case OP_LOGIT: st[sp - 1] = log(st[sp - 1] / (1.0 - st[sp - 1])); break;
            case OP_PHI: st[sp - 1] = 0.5 * erfc(-st[sp - 1] * 0.70710678118654752440); break;
            case OP_ICLOGLOG: st[sp - 1] = 1.0 - exp(-exp(st[sp - 1])); break;
            case OP_CLOGLOG: st[sp - 1] = log(-log(1.0 - st[sp - 1])); break;
            case OP_LOG1P: st[sp - 1] = log1p(st[sp - 1]); break;
            case OP_LGAMMA: st[sp - 1] = lgamma(st[sp - 1]); break;
            case OP_STEP: st[sp - 1] = st[sp - 1] >= 0 ? 1.0 : 0.0; break;
            case OP_CEIL: st[sp - 1] = ceil(st[sp - 1]); break;
            case OP_FLOOR: st[sp - 1] = floor(st[sp - 1]); break;
            case OP_ROUND: st[sp - 1] = nearbyint(st[sp - 1]); break;
            case OP_TRUNC: st[sp - 1] = trunc(st[sp - 1]); break;
            case OP_LT: st[sp - 2] = st[sp - 2] < st[sp - 1]; sp--; break;
            case OP_GT: st[sp - 2] = st[sp - 2] > st[sp - 1]; sp--; break;
            case OP_LE: st[sp - 2] = st[sp - 2] <= st[sp - 1]; sp--; break;
            case OP_GE: st[sp - 2] = st[sp - 2] >= st[sp - 1]; sp--; break;
            case OP_EQ: st[sp - 2] = st[sp - 2] == st[sp - 1]; sp--; break;
            case OP_NE: st[sp - 2] = st[sp - 2] != st[sp - 1]; sp--; break;
            case OP_OR: st[sp - 2] = (st[sp - 2] != 0.0) || (st[sp - 1] != 0.0); sp--; break;
            case OP_AND: st[sp - 2] = (st[sp - 2] != 0.0) && (st[sp - 1] != 0.0); sp--; break;
            case OP_NOT: st[sp - 1] = st[sp - 1] == 0.0; break;
            case OP_SEL: st[sp - 3] = st[sp - 3] != 0.0 ? st[sp - 2] : st[sp - 1]; sp -= 2; break;  /* cond, then, else */
            case OP_INPROD: {
                /* a = offset (in code[]) of two arg records (kind 1); other-dim indices were pushed A then B */
                const int* A = &code[a];
                const int* B = &code[a + ARGREC];
                double ob = 0, oa = 0;
                if (B[6] >= 0) ob = st[--sp];
                if (A[6] >= 0) oa = st[--sp];
                int n = A[5] - A[4] + 1;
                double s = 0;
                for (int e = 0; e < n; e++) s += slice_get(c, A, oa, e) * slice_get(c, B, ob, e);
                st[sp++] = s;
                break;
            }
            case OP_SUM: {
                const int* A = &code[a];
                double oa = 0;
                if (A[6] >= 0) oa = st[--sp];
                int n = A[5] - A[4] + 1;
                double s = 0;
                for (int e = 0; e < n; e++) s += slice_get(c, A, oa, e);
                st[sp++] = s;
                break;
            }
            default: return NAN;
        }
    }
}

static void set_loops(orc_ctx* c, const int* r, int inst) {
    for (int l = 0; l < r[R_NLOOP]; l++) {
        int n = r[R_HI0 + l] - r[R_LO0 + l] + 1;
        c->L[l] = r[R_LO0 + l] + inst % n;
        inst /= n;
    }
}

static int decl_ninst(const int* r) {
    int n = 1;
    for (int l = 0; l < r[R_NLOOP]; l++) n *= (r[R_HI0 + l] - r[R_LO0 + l] + 1);
    return n;
}

/* arena index of element e of the node's LHS (loops must be set) */
static int lhs_index(orc_ctx* c, const int* r, int e) {
    const orc_model* m = c->m;
    int var = r[R_VAR], off = m->var_off[var], nd = r[R_LNDIM];
    if (nd == 0) return off;
    int idx[2] = {0, 0};
    for (int k = 0; k < nd; k++) {
        if (r[R_LRDIM] == k)
            idx[k] = r[R_LRLO] + e;
        else
            idx[k] = (int)eval(c, r[R_LIDX0 + k]);
    }
    if (nd == 1) return off + idx[0] - 1;
    return off + (idx[0] - 1) + (idx[1] - 1) * m->var_nrow[var];
}
static int lhs_len(const int* r) { return r[R_LRDIM] >= 0 ? r[R_LRHI] - r[R_LRLO] + 1 : 1; }

/* (re)calculate one node; returns its logProb (0 for deterministic) and stores value / logProb */
static double calc_node(orc_ctx* c, int d, int inst) {
    const orc_model* m = c->m;
    const int* r = &m->decl[d * DREC];
    set_loops(c, r, inst);
    const int* a = &r[R_ARGS];
    if (r[R_KIND] == 0) {
        *vref(c, lhs_index(c, r, 0)) = eval(c, a[1]);
        return 0.0;
    }
    double lp;
    if (r[R_DIST] == D_MNORM_CHOL) {
        int n = lhs_len(r);
        double x[16], mu[16], ch[256], wk[16];
        const int* am = &a[0];
        const int* ac = &a[ARGREC];
        double other = am[6] >= 0 ? eval(c, am[6]) : 0;
        for (int e = 0; e < n; e++) {
            x[e] = *vref(c, lhs_index(c, r, e));
            mu[e] = slice_get(c, am, other, e);
        }
        int var = ac[1], nrow = m->var_nrow[var], off = m->var_off[var];
        for (int j = 0; j < n; j++)
            for (int i = 0; i < n; i++) ch[i + j * n] = *vref(c, off + (ac[2] - 1 + i) + (ac[4] - 1 + j) * nrow);
        int prec = (int)eval(c, a[2 * ARGREC + 1]);
        lp = apt_dmnorm_chol_log(x, mu, ch, n, prec, wk);
    } else if (r[R_DIST] == D_MULTI) { /* x[lo:hi] ~ dmulti(prob = vector slice, size = scalar) */
        int n = lhs_len(r);
        double x[64], pr[64];
        const int* ap = &a[0];
        double other = ap[6] >= 0 ? eval(c, ap[6]) : 0;
        for (int e = 0; e < n && e < 64; e++) {
            x[e] = *vref(c, lhs_index(c, r, e));
            pr[e] = slice_get(c, ap, other, e);
        }
        lp = apt_dmulti_log(x, pr, n, eval(c, a[ARGREC + 1]));
    } else if (r[R_DIST] == D_CAT) {
        double x = *vref(c, lhs_index(c, r, 0));
        const int* ap = &a[0];
        double other = ap[6] >= 0 ? eval(c, ap[6]) : 0;
        int n = ap[5] - ap[4] + 1;
        double pr[64];
        for (int e = 0; e < n && e < 64; e++) pr[e] = slice_get(c, ap, other, e);
        lp = apt_dcat_log(x, pr, n);
    } else {
        double x = *vref(c, lhs_index(c, r, 0));
        double p0 = r[R_NARGS] > 0 ? eval(c, a[1]) : 0;
        double p1 = r[R_NARGS] > 1 ? eval(c, a[ARGREC + 1]) : 0;
        switch (r[R_DIST]) {
            case D_NORM: lp = apt_dnorm_log(x, p0, p1); break;    /* (mean, sd) */
            case D_UNIF: lp = apt_dunif_log(x, p0, p1); break;    /* (min, max) */
            case D_GAMMA: lp = apt_dgamma_log(x, p0, p1); break;  /* (shape, scale) */
            case D_BETA: lp = apt_dbeta_log(x, p0, p1); break;    /* (shape1, shape2) */
            case D_BINOM: lp = apt_dbinom_log(x, p1, p0); break;  /* canonical (prob, size) */
            case D_POIS: lp = apt_dpois_log(x, p0); break;        /* (lambda) */
            case D_EXP: lp = apt_dexp_log(x, p0); break;          /* (scale) */
            default: lp = NAN;
        }
    }
    c->lp[m->lp_off[d] + inst] = lp;
    return lp;
}

/* nimble calculateDiff: for each node in order recompute; diff += new - stored; store new */
static double calculate_diff(orc_ctx* c, const int* nd, const int* ni, int n) {
    const orc_model* m = c->m;
    double diff = 0;
    for (int k = 0; k < n; k++) {
        int d = nd[k];
        if (m->decl[d * DREC + R_KIND] == 0) {
            calc_node(c, d, ni[k]);
        } else {
            double old = c->lp[m->lp_off[d] + ni[k]];
            double nw = calc_node(c, d, ni[k]);
            diff += nw - old;
        }
    }
    return diff;
}
static double get_logprob_nodes(orc_ctx* c, const int* nd, const int* ni, int n) {
    double s = 0;
    for (int k = 0; k < n; k++)
        if (c->m->decl[nd[k] * DREC + R_KIND] == 1) s += c->lp[c->m->lp_off[nd[k]] + ni[k]]; /* deterministic nodes carry no logProb */
    return s;
}
static double get_logprob_all(orc_ctx* c) {
    double s = 0;
    for (int i = 0; i < c->m->nlp; i++) s += c->lp[i];
    return s;
}
static void calculate_all(orc_ctx* c) {
    const orc_model* m = c->m;
    for (int d = 0; d < m->ndecl; d++) {
        int n = decl_ninst(&m->decl[d * DREC]);
        for (int i = 0; i < n; i++) calc_node(c, d, i);
    }
}

/* ------------------------------------------------------------------ construction */
orc_model* orc_model_new(int nvars, const int* var_off, const int* var_nrow, int nmut, int ntot, const double* arena0,
                         int ndecl, const int* decl, int ncode, const int* code, int nconst, const double* consts) {
    orc_model* m = (orc_model*)calloc(1, sizeof(orc_model));
    m->nvars = nvars; m->nmut = nmut; m->ntot = ntot; m->ndecl = ndecl; m->ncode = ncode; m->nconst = nconst;
    m->var_off = (int*)malloc(sizeof(int) * nvars); memcpy(m->var_off, var_off, sizeof(int) * nvars);
    m->var_nrow = (int*)malloc(sizeof(int) * nvars); memcpy(m->var_nrow, var_nrow, sizeof(int) * nvars);
    m->arena0 = (double*)malloc(sizeof(double) * ntot); memcpy(m->arena0, arena0, sizeof(double) * ntot);
    m->decl = (int*)malloc(sizeof(int) * ndecl * DREC); memcpy(m->decl, decl, sizeof(int) * ndecl * DREC);
    m->code = (int*)malloc(sizeof(int) * (ncode + 2)); memcpy(m->code, code, sizeof(int) * ncode);
    m->consts = (double*)malloc(sizeof(double) * (nconst + 1)); memcpy(m->consts, consts, sizeof(double) * nconst);
    m->lp_off = (int*)malloc(sizeof(int) * ndecl);
    int nlp = 0;
    for (int d = 0; d < ndecl; d++) {
        m->lp_off[d] = nlp;
        if (m->decl[d * DREC + R_KIND] == 1) nlp += decl_ninst(&m->decl[d * DREC]);
    }
    m->nlp = nlp;
    return m;
}
void orc_model_free(orc_model* m) {
    if (!m) return;
    free(m->var_off); free(m->var_nrow); free(m->arena0); free(m->decl); free(m->code); free(m->consts); free(m->lp_off);
    free(m);
}
int orc_model_nlp(const orc_model* m) { return m->nlp; }

/* whole-model log probability of an arbitrary state of the mutable arena prefix (PARAM values are read from
   `mut`; deterministic nodes and all logProbs are recomputed).  out_lp_decl (ndecl) receives per-declaration sums. */
double orc_model_logprob(const orc_model* m, const double* mut_in, double* out_lp_decl) {
    orc_ctx c;
    c.m = m;
    c.mut = (double*)malloc(sizeof(double) * (m->nmut + 1));
    c.lp = (double*)calloc(m->nlp + 1, sizeof(double));
    memcpy(c.mut, mut_in, sizeof(double) * m->nmut);
    calculate_all(&c);
    double tot = get_logprob_all(&c);
    if (out_lp_decl)
        for (int d = 0; d < m->ndecl; d++) {
            double s = 0;
            if (m->decl[d * DREC + R_KIND] == 1) {
                int n = decl_ninst(&m->decl[d * DREC]);
                for (int i = 0; i < n; i++) s += c.lp[m->lp_off[d] + i];
            }
            out_lp_decl[d] = s;
        }
    free(c.mut); free(c.lp);
    return tot;
}

/* The same node log-probabilities, summed in extended precision (long double, 64-bit mantissa).  model$getLogProb()
 * adds the nodes one after the other in double; over 1e6 nodes of nearly equal size that sequential sum itself is off by
 * ~1e-11 relative (each addition rounds a sum ~1e6 times larger than the addend), which is more than the 1e-12 the north
 * star asks of logProb agreement.  This entry point gives the value those node terms really add up to; the parity test
 * of config 3 at N = 1e6 compares the device against it and reports the distance of the literal sequential sum. */
double orc_model_logprob_ld(const orc_model* m, const double* mut_in, double* out_lp_decl) {
    orc_ctx c;
    c.m = m;
    c.mut = (double*)malloc(sizeof(double) * (m->nmut + 1));
    c.lp = (double*)calloc(m->nlp + 1, sizeof(double));
    memcpy(c.mut, mut_in, sizeof(double) * m->nmut);
    calculate_all(&c);
    long double tot = 0.0L;
    for (int d = 0; d < m->ndecl; d++) {
        long double s = 0.0L;
        if (m->decl[d * DREC + R_KIND] == 1) {
            int n = decl_ninst(&m->decl[d * DREC]);
            for (int i = 0; i < n; i++) s += (long double)c.lp[m->lp_off[d] + i];
        }
        if (out_lp_decl) out_lp_decl[d] = (double)s;
        tot += s;
    }
    free(c.mut); free(c.lp);
    return (double)tot;
}

orc_sspec* orc_specs_new(int ns) { return (orc_sspec*)calloc(ns, sizeof(orc_sspec)); }

int orc_spec_set(const orc_model* m, orc_sspec* specs, int s, int kind, int d, const int* tgt, int ncalc,
                 const int* calc_decl, const int* calc_inst, int ntnode, const int* tn_decl, const int* tn_inst,
                 int log_scale, int reflective, int adaptive, int adapt_scale_only, int adapt_interval,
                 int temper_priors, double adapt_exp, double scale0, const double* propcov0, int lower_code,
                 int upper_code) {
    orc_sspec* p = &specs[s];
    p->kind = kind; p->d = d;
    p->tgt = (int*)malloc(sizeof(int) * d); memcpy(p->tgt, tgt, sizeof(int) * d);
    p->ncalc = ncalc;
    p->calc_decl = (int*)malloc(sizeof(int) * (ncalc + 1)); memcpy(p->calc_decl, calc_decl, sizeof(int) * ncalc);
    p->calc_inst = (int*)malloc(sizeof(int) * (ncalc + 1)); memcpy(p->calc_inst, calc_inst, sizeof(int) * ncalc);
    p->ntnode = ntnode;
    p->tn_decl = (int*)malloc(sizeof(int) * (ntnode + 1)); memcpy(p->tn_decl, tn_decl, sizeof(int) * ntnode);
    p->tn_inst = (int*)malloc(sizeof(int) * (ntnode + 1)); memcpy(p->tn_inst, tn_inst, sizeof(int) * ntnode);
    p->log_scale = log_scale; p->reflective = reflective; p->adaptive = adaptive;
    p->adapt_scale_only = adapt_scale_only; p->adapt_interval = adapt_interval; p->temper_priors = temper_priors;
    p->adapt_exp = adapt_exp; p->scale0 = scale0;
    p->propcov0 = (double*)malloc(sizeof(double) * d * d);
    if (propcov0) memcpy(p->propcov0, propcov0, sizeof(double) * d * d);
    else for (int i = 0; i < d * d; i++) p->propcov0[i] = (i % (d + 1) == 0) ? 1.0 : 0.0;
    p->lower_code = lower_code; p->upper_code = upper_code;
    p->max_steps = 100; p->discrete = 0;
    /* copy lists for nimCopy(nodes = calcNodes, logProb = TRUE)  (APT:722-724, decideAndJump) */
    orc_ctx c; c.m = m; c.mut = NULL; c.lp = NULL;
    int nv = 0, nl = 0;
    for (int k = 0; k < ncalc; k++) {
        const int* r = &m->decl[calc_decl[k] * DREC];
        nv += lhs_len(r);
        if (r[R_KIND] == 1) nl++;
    }
    p->copyv = (int*)malloc(sizeof(int) * (nv + 1));
    p->copylp = (int*)malloc(sizeof(int) * (nl + 1));
    nv = nl = 0;
    double dummy[1] = {0};
    c.mut = dummy; /* lhs index expressions only read loop registers / constants */
    for (int k = 0; k < ncalc; k++) {
        const int* r = &m->decl[calc_decl[k] * DREC];
        set_loops(&c, r, calc_inst[k]);
        int len = lhs_len(r);
        for (int e = 0; e < len; e++) {
            int idx = lhs_index(&c, r, e);
            if (idx < m->nmut) p->copyv[nv++] = idx;
        }
        if (r[R_KIND] == 1) p->copylp[nl++] = m->lp_off[calc_decl[k]] + calc_inst[k];
    }
    p->ncopyv = nv; p->ncopylp = nl;
    return 0;
}

/* sampler_RW_multinomial_tempered (kind 3): temper_priors of orc_spec_set carries control$useTempering */
int orc_spec_set_multinomial(orc_sspec* specs, int s, double ntotal) {
    specs[s].ntotal = ntotal;
    return 0;
}

/* extra control of sampler_slice_tempered (kind 2); scale0 of orc_spec_set carries control$width */
int orc_spec_set_slice(orc_sspec* specs, int s, int max_steps, int discrete) {
    specs[s].max_steps = max_steps; specs[s].discrete = discrete;
    return 0;
}

static void sstate_reset(const orc_sspec* p, orc_sstate* s) {
    s->scale = p->scale0; s->gamma1 = 0; s->sumJumps = 0; s->timesRan = s->timesAccepted = s->timesAdapted = 0;
    if (p->kind == 1) {
        int d = p->d;
        memcpy(s->propCov, p->propcov0, sizeof(double) * d * d);
        apt_chol_upper(s->propCov, s->chol, d);
        for (int i = 0; i < d * d; i++) s->cholScaled[i] = s->chol[i] * s->scale;
    }
    if (p->kind == 3) { /* reset (APT:1133-1142): u is NOT reset */
        int dd = p->d * p->d;
        for (int i = 0; i < dd; i++) {
            s->mn[i] = 0; s->mn[dd + i] = 0; s->mn[2 * dd + i] = 0; s->mn[3 * dd + i] = 0; s->mn[4 * dd + i] = 0;
            s->mn[5 * dd + i] = 1; s->mn[6 * dd + i] = 1; s->mn[7 * dd + i] = 0.2;
        }
    }
}

orc_ladder* orc_ladder_new(const orc_model* m, const orc_sspec* specs, int NS, const double* temps, int K,
                           uint32_t seed, uint32_t ladder, double ULT, int monitorTmax, int nmon, const int* mon,
                           int nmon2, const int* mon2) {
    orc_ladder* L = (orc_ladder*)calloc(1, sizeof(orc_ladder));
    L->m = m; L->K = K; L->NS = NS; L->specs = specs; L->seed = seed; L->ladder = ladder; L->ULT = ULT;
    L->monitorTmax = monitorTmax;
    L->nmon = nmon; L->mon = (int*)malloc(sizeof(int) * (nmon + 1)); memcpy(L->mon, mon, sizeof(int) * nmon);
    L->nmon2 = nmon2; L->mon2 = (int*)malloc(sizeof(int) * (nmon2 + 1)); memcpy(L->mon2, mon2, sizeof(int) * nmon2);
    L->Temps = (double*)malloc(sizeof(double) * K);
    L->invTemps = (double*)malloc(sizeof(double) * K);
    L->TempsCurrent = (double*)malloc(sizeof(double) * K);
    L->logProbTemps = (double*)calloc(K, sizeof(double));
    L->pSwap = (double*)calloc((size_t)K * K, sizeof(double));
    L->temporary = (double*)calloc(K, sizeof(double));
    L->accCountSwap = (double*)calloc((size_t)K * K, sizeof(double));
    /* Temps <- sort(Temps)  (APT:266) */
    memcpy(L->Temps, temps, sizeof(double) * K);
    for (int i = 1; i < K; i++) {
        double t = L->Temps[i]; int j = i - 1;
        while (j >= 0 && L->Temps[j] > t) { L->Temps[j + 1] = L->Temps[j]; j--; }
        L->Temps[j + 1] = t;
    }
    for (int i = 0; i < K; i++) { L->invTemps[i] = 1.0 / L->Temps[i]; L->TempsCurrent[i] = L->Temps[i]; }
    L->rung = (orc_buf*)calloc(K, sizeof(orc_buf));
    L->saved = (orc_buf*)calloc(K, sizeof(orc_buf));
    for (int k = 0; k < K; k++) {
        L->rung[k].mut = (double*)malloc(sizeof(double) * (m->nmut + 1));
        L->rung[k].lp = (double*)calloc(m->nlp + 1, sizeof(double));
        L->saved[k].mut = (double*)malloc(sizeof(double) * (m->nmut + 1));
        L->saved[k].lp = (double*)calloc(m->nlp + 1, sizeof(double));
    }
    L->model.mut = (double*)malloc(sizeof(double) * (m->nmut + 1));
    L->model.lp = (double*)calloc(m->nlp + 1, sizeof(double));
    memcpy(L->model.mut, m->arena0, sizeof(double) * m->nmut);
    L->ss = (orc_sstate*)calloc((size_t)K * NS, sizeof(orc_sstate));
    for (int k = 0; k < K; k++)
        for (int s = 0; s < NS; s++) {
            orc_sstate* st = &L->ss[s + k * NS];
            const orc_sspec* p = &specs[s];
            if (p->kind == 1) {
                int d = p->d;
                st->propCov = (double*)malloc(sizeof(double) * d * d);
                st->chol = (double*)malloc(sizeof(double) * d * d);
                st->cholScaled = (double*)malloc(sizeof(double) * d * d);
                st->empir = (double*)calloc((size_t)p->adapt_interval * d, sizeof(double));
            }
            if (p->kind == 3) { st->mn = (double*)calloc((size_t)8 * p->d * p->d, sizeof(double)); st->u = -1.0; }
            sstate_reset(p, st);
        }
    L->totalIters = 1;   /* APT:321 */
    L->resetAnyway = 1;  /* APT:320 */
    L->rngIter = 0;
    return L;
}

void orc_ladder_free(orc_ladder* L) {
    if (!L) return;
    for (int k = 0; k < L->K; k++) { free(L->rung[k].mut); free(L->rung[k].lp); free(L->saved[k].mut); free(L->saved[k].lp); }
    for (int i = 0; i < L->K * L->NS; i++) { free(L->ss[i].propCov); free(L->ss[i].chol); free(L->ss[i].cholScaled); free(L->ss[i].empir); free(L->ss[i].mn); }
    free(L->rung); free(L->saved); free(L->model.mut); free(L->model.lp); free(L->ss);
    free(L->Temps); free(L->invTemps); free(L->TempsCurrent); free(L->logProbTemps); free(L->pSwap);
    free(L->temporary); free(L->accCountSwap); free(L->mon); free(L->mon2);
    free(L->samples); free(L->samples2); free(L->samplesTmax); free(L->samples2Tmax); free(L->logProbs); free(L->tempTraj);
    free(L);
}

void orc_ladder_set_tables(orc_ladder* L, const double* tabZ, const double* tabU, const double* tabS, int dmax,
                           unsigned char* recDec, int* recSwap) {
    L->tabZ = tabZ; L->tabU = tabU; L->tabS = tabS; L->tabDmax = dmax; L->recDec = recDec; L->recSwap = recSwap;
}
/* set the values of the model's mutable arena prefix (inits) */
void orc_ladder_set_model(orc_ladder* L, const double* mut) { memcpy(L->model.mut, mut, sizeof(double) * L->m->nmut); }

/* ------------------------------------------------------------------ randoms */
static void draw_normals(orc_ladder* L, int s, int tt, int d, double* z) {
    if (L->tabZ) {
        const double* p = L->tabZ + (((size_t)L->itRun * L->K + tt) * L->NS + s) * L->tabDmax;
        for (int e = 0; e < d; e++) z[e] = p[e];
        return;
    }
    for (int b = 0; 2 * b < d; b++) {
        double z0, z1;
        apt_slot_normal2(L->seed, L->ladder, L->rngIter, (uint32_t)s, (uint32_t)tt, (uint32_t)b, &z0, &z1);
        z[2 * b] = z0;
        if (2 * b + 1 < d) z[2 * b + 1] = z1;
    }
}
static double draw_accept_u(orc_ladder* L, int s, int tt) {
    if (L->tabU) return L->tabU[((size_t)L->itRun * L->K + tt) * L->NS + s];
    return apt_slot_accept_u(L->seed, L->ladder, L->rngIter, (uint32_t)s, (uint32_t)tt);
}
static void draw_swap_u(orc_ladder* L, int ii, double* ucat, double* uacc) {
    if (L->tabS) {
        const double* p = L->tabS + ((size_t)L->itRun * L->K + ii) * 2;
        *ucat = p[0]; *uacc = p[1];
        return;
    }
    apt_slot_swap_u(L->seed, L->ladder, L->rngIter, (uint32_t)ii, ucat, uacc);
}

/* nimble decide(logMHR): NaN -> FALSE; > 0 -> TRUE without a draw; else runif < exp(logMHR).
   With slot-addressed randoms the uniform's slot is fixed whether or not it is consumed. */
static int decide(double logMHR, double u) {
    if (isnan(logMHR)) return 0;
    if (logMHR > 0) return 1;
    return u < exp(logMHR);
}

/* ------------------------------------------------------------------ samplers */
static void copy_nodes(const orc_sspec* p, const orc_buf* from, orc_buf* to) {
    for (int i = 0; i < p->ncopyv; i++) to->mut[p->copyv[i]] = from->mut[p->copyv[i]];
    for (int i = 0; i < p->ncopylp; i++) to->lp[p->copylp[i]] = from->lp[p->copylp[i]];
}

/* sampler_RW_tempered$run  (APT:690-729) + adaptiveProcedure (APT:734-747) */
static int run_scalar(orc_ladder* L, int s, int tt, orc_buf* model, orc_buf* saved, double temperature) {
    const orc_sspec* p = &L->specs[s];
    orc_sstate* st = &L->ss[s + tt * L->NS];
    orc_ctx c; c.m = L->m; c.mut = model->mut; c.lp = model->lp;
    double z;
    draw_normals(L, s, tt, 1, &z);
    double currentValue = model->mut[p->tgt[0]];
    double propLogScale = 0, propValue;
    if (p->log_scale) {
        propLogScale = 0 + st->scale * z; /* rnorm(1, mean=0, sd=scale) */
        propValue = currentValue * exp(propLogScale);
    } else {
        propValue = currentValue + st->scale * z; /* rnorm(1, mean=currentValue, sd=scale) */
    }
    if (p->reflective) {
        double lower = p->lower_code >= 0 ? eval(&c, p->lower_code) : -INFINITY;
        double upper = p->upper_code >= 0 ? eval(&c, p->upper_code) : INFINITY;
        while (propValue < lower || propValue > upper) {
            if (propValue < lower) propValue = 2 * lower - propValue;
            if (propValue > upper) propValue = 2 * upper - propValue;
        }
    }
    double logMHR;
    if (p->temper_priors) {
        model->mut[p->tgt[0]] = propValue;
        logMHR = calculate_diff(&c, p->calc_decl, p->calc_inst, p->ncalc) / temperature + propLogScale;
    } else {
        double logPriorWeightOri = get_logprob_nodes(&c, p->tn_decl, p->tn_inst, p->ntnode);
        model->mut[p->tgt[0]] = propValue;
        double diffLogProb = calculate_diff(&c, p->calc_decl, p->calc_inst, p->ncalc);
        double logPriorWeightProp = get_logprob_nodes(&c, p->tn_decl, p->tn_inst, p->ntnode);
        double diffLogPriorWeights = logPriorWeightProp - logPriorWeightOri;
        logMHR = (diffLogProb - diffLogPriorWeights) / temperature + diffLogPriorWeights + propLogScale;
    }
    int jump = decide(logMHR, draw_accept_u(L, s, tt));
    if (jump) copy_nodes(p, model, saved); else copy_nodes(p, saved, model);
    if (p->adaptive) {
        st->timesRan++;
        if (jump) st->timesAccepted++;
        if (st->timesRan % p->adapt_interval == 0) {
            double acceptanceRate = (double)st->timesAccepted / st->timesRan;
            st->timesAdapted++;
            st->gamma1 = 1 / pow(st->timesAdapted + 3, p->adapt_exp);
            double gamma2 = 10 * st->gamma1;
            double adaptFactor = exp(gamma2 * (acceptanceRate - 0.44));
            st->scale = st->scale * adaptFactor;
            st->timesRan = 0;
            st->timesAccepted = 0;
        }
    }
    return jump;
}

/* sampler_RW_block_tempered$run (APT:821-837), generateProposalVector (APT:842-846),
   adaptiveProcedure (APT:847-872), nimble calcAdaptationFactor */
static int run_block(orc_ladder* L, int s, int tt, orc_buf* model, orc_buf* saved, double temperature) {
    const orc_sspec* p = &L->specs[s];
    orc_sstate* st = &L->ss[s + tt * L->NS];
    orc_ctx c; c.m = L->m; c.mut = model->mut; c.lp = model->lp;
    int d = p->d;
    double* z = (double*)alloca(sizeof(double) * (d + 1));
    double* prop = (double*)alloca(sizeof(double) * d);
    draw_normals(L, s, tt, d, z);
    /* rmnorm_chol(mean, chol = U, prec_param = 0): x = mean + U' z */
    for (int j = 0; j < d; j++) {
        double acc = 0;
        for (int i = 0; i <= j; i++) acc += st->cholScaled[i + j * d] * z[i];
        prop[j] = model->mut[p->tgt[j]] + acc;
    }
    double lpMHR;
    if (p->temper_priors) {
        for (int j = 0; j < d; j++) model->mut[p->tgt[j]] = prop[j];
        lpMHR = calculate_diff(&c, p->calc_decl, p->calc_inst, p->ncalc) / temperature;
    } else {
        double logPriorWeightOri = get_logprob_nodes(&c, p->tn_decl, p->tn_inst, p->ntnode);
        for (int j = 0; j < d; j++) model->mut[p->tgt[j]] = prop[j];
        lpMHR = calculate_diff(&c, p->calc_decl, p->calc_inst, p->ncalc);
        double diffLogPriorWeights = get_logprob_nodes(&c, p->tn_decl, p->tn_inst, p->ntnode) - logPriorWeightOri;
        lpMHR = (lpMHR - diffLogPriorWeights) / temperature + diffLogPriorWeights;
    }
    int jump = decide(lpMHR, draw_accept_u(L, s, tt));
    if (jump) copy_nodes(p, model, saved); else copy_nodes(p, saved, model);
    if (p->adaptive) {
        int AI = p->adapt_interval;
        st->timesRan++;
        if (jump) st->timesAccepted++;
        if (!p->adapt_scale_only)
            for (int j = 0; j < d; j++) st->empir[(st->timesRan - 1) + (size_t)j * AI] = model->mut[p->tgt[j]];
        if (st->timesRan % AI == 0) {
            static const double optAR[5] = {0.44, 0.35, 0.32, 0.25, 0.234};
            double acceptanceRate = (double)st->timesAccepted / st->timesRan;
            st->timesAdapted++;
            st->gamma1 = 1 / pow(st->timesAdapted + 3, p->adapt_exp);
            double gamma2 = 10 * st->gamma1;
            double adaptFactor = exp(gamma2 * (acceptanceRate - optAR[(d < 5 ? d : 5) - 1]));
            st->scale = st->scale * adaptFactor;
            if (!p->adapt_scale_only) {
                double gamma1 = st->gamma1;
                for (int j = 0; j < d; j++) {
                    double mean = 0;
                    for (int t = 0; t < AI; t++) mean += st->empir[t + (size_t)j * AI];
                    mean /= AI;
                    for (int t = 0; t < AI; t++) st->empir[t + (size_t)j * AI] -= mean;
                }
                for (int a = 0; a < d; a++)
                    for (int b = 0; b < d; b++) {
                        double sum = 0;
                        for (int t = 0; t < AI; t++) sum += st->empir[t + (size_t)a * AI] * st->empir[t + (size_t)b * AI];
                        double empirCov = sum / (st->timesRan - 1);
                        st->propCov[a + b * d] = st->propCov[a + b * d] + gamma1 * (empirCov - st->propCov[a + b * d]);
                    }
                apt_chol_upper(st->propCov, st->chol, d);
            }
            for (int i = 0; i < d * d; i++) st->cholScaled[i] = st->chol[i] * st->scale;
            st->timesRan = 0;
            st->timesAccepted = 0;
        }
    }
    return jump;
}

/* sampler_slice_tempered$run (APT:929-964), setAndCalculateTarget (APT:969-975), adaptiveProcedure (APT:976-987).
   Uniform number j of the update (in the order the reference draws them: rexp, L offset, maxStepsL, x1, shrinkage
   draws ...) is read from slot block j / 2, words (0,1) for even j and (2,3) for odd j; rexp(1) = -log(U).
   Returns the number of setAndCalculateTarget calls (recorded modulo 256 as the unit's "decision"). */
typedef struct { orc_ladder* L; int s, tt; uint32_t j; uint32_t w[4]; } slice_rng;
static double slice_unif(slice_rng* r) {
    if ((r->j & 1u) == 0) apt_slot_words(r->L->seed, r->L->ladder, r->L->rngIter, (uint32_t)r->s, (uint32_t)r->tt, r->j >> 1, r->w);
    double u = (r->j & 1u) ? apt_u53(r->w[2], r->w[3]) : apt_u53(r->w[0], r->w[1]);
    r->j++;
    return u;
}
static double slice_set_and_calc(orc_ctx* c, const orc_sspec* p, orc_buf* model, double value, int* ncalls) {
    if (p->discrete) value = floor(value);
    model->mut[p->tgt[0]] = value;
    double lp = 0;
    for (int k = 0; k < p->ncalc; k++) lp += calc_node(c, p->calc_decl[k], p->calc_inst[k]); /* calculate(model, calcNodes) */
    (*ncalls)++;
    return lp;
}
static int run_slice(orc_ladder* L, int s, int tt, orc_buf* model, orc_buf* saved, double temperature) {
    const orc_sspec* p = &L->specs[s];
    orc_sstate* st = &L->ss[s + tt * L->NS];
    orc_ctx c; c.m = L->m; c.mut = model->mut; c.lp = model->lp;
    slice_rng rg; rg.L = L; rg.s = s; rg.tt = tt; rg.j = 0;
    int ncalls = 0;
    double width = st->scale;
    double u = get_logprob_nodes(&c, p->calc_decl, p->calc_inst, p->ncalc) / temperature - (-log(slice_unif(&rg)));
    double x0 = model->mut[p->tgt[0]];
    double Lb = x0 - slice_unif(&rg) * width;
    double Rb = Lb + width;
    int maxStepsL = (int)floor(slice_unif(&rg) * p->max_steps);
    int maxStepsR = p->max_steps - 1 - maxStepsL;
    double lp = slice_set_and_calc(&c, p, model, Lb, &ncalls) / temperature;
    while (maxStepsL > 0 && !isnan(lp) && lp >= u) {
        Lb = Lb - width;
        lp = slice_set_and_calc(&c, p, model, Lb, &ncalls) / temperature;
        maxStepsL--;
    }
    lp = slice_set_and_calc(&c, p, model, Rb, &ncalls) / temperature;
    while (maxStepsR > 0 && !isnan(lp) && lp >= u) {
        Rb = Rb + width;
        lp = slice_set_and_calc(&c, p, model, Rb, &ncalls) / temperature;
        maxStepsR--;
    }
    double x1 = Lb + slice_unif(&rg) * (Rb - Lb);
    lp = slice_set_and_calc(&c, p, model, x1, &ncalls) / temperature;
    while (isnan(lp) || lp < u) {
        if (x1 < x0) Lb = x1; else Rb = x1;
        x1 = Lb + slice_unif(&rg) * (Rb - Lb);
        lp = slice_set_and_calc(&c, p, model, x1, &ncalls) / temperature;
    }
    copy_nodes(p, model, saved);
    double jumpDist = fabs(x1 - x0);
    if (p->adaptive) {
        st->timesRan++;
        st->sumJumps += jumpDist;
        if (st->timesRan % p->adapt_interval == 0) {
            double adaptFactor = pow(3.0 / 4.0, st->timesAdapted);
            double meanJump = st->sumJumps / st->timesRan;
            st->scale = st->scale + (2 * meanJump - st->scale) * adaptFactor;
            st->timesAdapted++;
            st->timesRan = 0;
            st->sumJumps = 0;
        }
    }
    return ncalls & 0xFF;
}

/* sampler_RW_multinomial_tempered$run (APT:1054-1082), generateProposalVector (APT:1087-1098),
   adaptiveProcedure (APT:1099-1129).  Randoms of sub-proposal q (0-based, in loop order): slot block q, words (0,1) ->
   the uniform inverted by rbinom, words (2,3) -> the acceptance uniform; the recycled u (APT:1040) is drawn as
   pi * U(words 0,1 of block 0x7FFF) at the sampler's first update.  Returns the number of accepted sub-proposals. */
#define APT_SLOT_MULTI_U0 0x7FFFu
static int g_multi_exact_reverse = 0;
void orc_set_multinomial_exact_reverse(int on) { g_multi_exact_reverse = on; }
static int run_multinomial(orc_ladder* L, int s, int tt, orc_buf* model, orc_buf* saved, double temperature) {
    const orc_sspec* p = &L->specs[s];
    orc_sstate* st = &L->ss[s + tt * L->NS];
    orc_ctx c; c.m = L->m; c.mut = model->mut; c.lp = model->lp;
    const int d = p->d, dd = d * d;
    double *timesRan = st->mn, *AcceptRates = st->mn + dd, *ScaleShifts = st->mn + 2 * dd, *totalAdapted = st->mn + 3 * dd,
           *timesAccepted = st->mn + 4 * dd, *EN = st->mn + 5 * dd, *ENDelta = st->mn + 6 * dd, *Rescale = st->mn + 7 * dd;
    const double Pi = 3.14159265358979323846, PiOver2 = Pi / 2, Ntotal = p->ntotal, NOverL = Ntotal / d;
    uint32_t w[4];
    if (st->u < 0) { /* u <- runif(1, 0, Pi)  (APT:1040) */
        apt_slot_words(L->seed, L->ladder, L->rngIter, (uint32_t)s, (uint32_t)tt, APT_SLOT_MULTI_U0, w);
        st->u = Pi * apt_u53(w[0], w[1]);
    }
    double u = st->u;
    double prop[64];
    int q = 0, naccept = 0;
    for (int iFROM = 1; iFROM <= d; iFROM++)
        for (int iTO = 1; iTO <= d - 1; iTO++, q++) {
            int iFrom, iTo;
            if (u > PiOver2) {
                iFrom = iFROM; iTo = iTO;
                if (iFrom == iTo) iTo = d;
                u = 2 * (u - PiOver2);
            } else {
                iFrom = iTO; iTo = iFROM;
                if (iFrom == iTo) iFrom = d;
                u = 2 * (PiOver2 - u);
            }
            const int ft = (iFrom - 1) + (iTo - 1) * d, tf = (iTo - 1) + (iFrom - 1) * d;
            apt_slot_words(L->seed, L->ladder, L->rngIter, (uint32_t)s, (uint32_t)tt, (uint32_t)q, w);
            /* generateProposalVector(iFrom, iTo) */
            for (int e = 0; e < d; e++) prop[e] = model->mut[p->tgt[e]];
            double pSwap = fmin(1, fmax(1, EN[ft]) / prop[iFrom - 1]);
            double nSwap = apt_rbinom_inv(prop[iFrom - 1], pSwap, apt_u53(w[0], w[1]));
            double lpProp = apt_dbinom_log(nSwap, prop[iFrom - 1], pSwap);
            prop[iFrom - 1] -= nSwap;
            prop[iTo - 1] += nSwap;
            /* APT:1094 divides by propVector[iTo] + nSwap AFTER propVector[iTo] was incremented (APT:1093), i.e. by
               x_to + 2 nSwap, while lpRev (APT:1095) is evaluated at size x_to + nSwap: the proposal ratio is not the
               exact reverse probability and the chain's stationary law deviates slightly from the posterior.  Restated
               as the reference has it; the test-only switch below evaluates the exact reverse proposal instead, which
               pins the rest of the restatement against an enumerated posterior (tests/test_oracle_engine.py). */
            double pRevSwap = fmin(1, fmax(1, EN[tf]) / (prop[iTo - 1] + (g_multi_exact_reverse ? 0.0 : nSwap)));
            double lpRev = apt_dbinom_log(nSwap, prop[iTo - 1], pRevSwap);
            /* my_setAndCalculateDiff$run(propValueVector) */
            for (int e = 0; e < d; e++) model->mut[p->tgt[e]] = prop[e];
            double diff = calculate_diff(&c, p->calc_decl, p->calc_inst, p->ncalc);
            double lpMHR = p->temper_priors ? diff / temperature + lpRev - lpProp : diff + lpRev - lpProp;
            int jump = decide(lpMHR, apt_u53(w[2], w[3])); /* my_decideAndJump$run(lpMHR, 0, 0, 0) */
            if (jump) copy_nodes(p, model, saved); else copy_nodes(p, saved, model);
            naccept += jump;
            if (p->adaptive) { /* adaptiveProcedure(jump, iFrom, iTo) */
                timesRan[ft] += 1;
                if (jump) timesAccepted[ft] += 1;
                if (fmod(timesRan[ft], (double)p->adapt_interval) == 0) {
                    totalAdapted[ft] += 1;
                    double accRate = timesAccepted[ft] / timesRan[ft];
                    AcceptRates[ft] = accRate;
                    if (accRate > 0.5) EN[ft] = fmin(Ntotal, EN[ft] + ENDelta[ft] / totalAdapted[ft]);
                    else EN[ft] = fmax(1, EN[ft] - ENDelta[ft] / totalAdapted[ft]);
                    if (accRate < Rescale[ft] || accRate > (1 - Rescale[ft])) {
                        if (EN[ft] > 1 && EN[ft] < Ntotal) {
                            ScaleShifts[ft] += 1;
                            ENDelta[ft] = fmin(NOverL, ENDelta[ft] * totalAdapted[ft] / 10);
                            ENDelta[tf] = ENDelta[ft];
                            Rescale[ft] = 0.2 * pow(0.95, ScaleShifts[ft]);
                        }
                    }
                    if (EN[ft] < 1) EN[ft] = 1;
                    EN[tf] = EN[ft];
                    timesRan[ft] = 0;
                    timesAccepted[ft] = 0;
                }
            }
        }
    st->u = u;
    return naccept & 0xFF;
}

/* ------------------------------------------------------------------ swap machinery */
/* pSwapCalc  (APT:556-576) */
static void pswap_calc(orc_ladder* L) {
    int K = L->K;
    double* P = L->pSwap; /* P[i + j*K] */
    for (int ii = 1; ii < K; ii++)
        for (int jj = 0; jj < ii; jj++) {
            P[ii + jj * K] = exp(-fabs(L->logProbTemps[ii] - L->logProbTemps[jj]));
            P[jj + ii * K] = P[ii + jj * K];
        }
    for (int ii = 0; ii < K; ii++) {
        double rowSum = 0;
        for (int jj = 0; jj < K; jj++) rowSum += P[ii + jj * K];
        if (rowSum > 0) {
            for (int jj = 0; jj < K; jj++) P[ii + jj * K] = P[ii + jj * K] / rowSum;
        } else {
            for (int jj = 0; jj < K; jj++) P[ii + jj * K] = (double)(ii != jj) / (K - 1);
        }
    }
}
/* nimble rcat(1, prob) [NIMBLE-UNVERIFIED]: one uniform, inverse CDF over prob/sum(prob) */
static int rcat(const double* prob, int stride, int K, double u) {
    double sum = 0;
    for (int j = 0; j < K; j++) sum += prob[j * stride];
    double target = u * sum, cdf = 0;
    for (int j = 0; j < K; j++) {
        cdf += prob[j * stride];
        if (target < cdf) return j;
    }
    for (int j = K - 1; j >= 0; j--) if (prob[j * stride] > 0) return j;
    return K - 1;
}

static void grow(double** p, long rows, int cols) { *p = (double*)realloc(*p, sizeof(double) * (size_t)(rows > 0 ? rows : 1) * (cols > 0 ? cols : 1)); }

static void write_row(const orc_ladder* L, const orc_buf* b, const int* mon, int nmon, double* dst) {
    for (int c = 0; c < nmon; c++) dst[c] = mon[c] >= 0 ? b->mut[mon[c]] : b->lp[-mon[c] - 1];
}

/* buildAPT$run  (APT:336-549) */
int orc_ladder_run(orc_ladder* L, long niter, int reset, int resetMV, int resetTempering, int adaptTemps,
                   double tuneTemper1, double tuneTemper2, int thin, int thin2) {
    const orc_model* m = L->m;
    int K = L->K, NS = L->NS;
    if (niter < 0) return -1;
    if (thin < 1) thin = 1;
    if (thin2 < 1) thin2 = 1;
    if (L->resetAnyway) { /* APT:358-363 */
        reset = 1; resetTempering = 1; L->resetAnyway = 0; L->totalIters = 1;
    }
    { /* my_initializeModel$run()  (APT:364): deterministic nodes + all logProbs recomputed */
        orc_ctx c; c.m = m; c.mut = L->model.mut; c.lp = L->model.lp;
        calculate_all(&c);
    }
    if (resetTempering) L->totalIters = 1;
    memset(L->accCountSwap, 0, sizeof(double) * K * K);
    long off, off2;
    if (reset) {
        for (int tt = 0; tt < K; tt++) {
            memcpy(L->rung[tt].mut, L->model.mut, sizeof(double) * m->nmut);
            memcpy(L->rung[tt].lp, L->model.lp, sizeof(double) * m->nlp);
        }
        for (int tt = 0; tt < K; tt++) for (int s = 0; s < NS; s++) sstate_reset(&L->specs[s], &L->ss[s + tt * NS]);
        off = 0; off2 = 0;
    } else {
        off = L->nrows; off2 = L->nrows2;
        if (resetMV) { off = 0; off2 = 0; }
    }
    long nr = niter / thin, nr2 = niter / thin2;
    L->ttrows = nr;
    grow(&L->tempTraj, nr, K);
    grow(&L->samples, off + nr, L->nmon); grow(&L->samples2, off2 + nr2, L->nmon2);
    grow(&L->samplesTmax, off + nr, L->nmon); grow(&L->samples2Tmax, off2 + nr2, L->nmon2);
    grow(&L->logProbs, off + nr, 1);
    L->nrows = off + nr; L->nrows2 = off2 + nr2;
    /* per-rung mvSaved starts as a copy of the rung (APT:439 does this copy every iteration) */
    for (int tt = 0; tt < K; tt++) {
        memcpy(L->saved[tt].mut, L->rung[tt].mut, sizeof(double) * m->nmut);
        memcpy(L->saved[tt].lp, L->rung[tt].lp, sizeof(double) * m->nlp);
    }
    for (long iter = 1; iter <= niter; iter++) {
        L->itRun = iter - 1;
        /* ---- random-walk phase (APT:435-448) */
        for (int tt = 0; tt < K; tt++) {
            orc_buf* model = &L->rung[tt];
            orc_buf* saved = &L->saved[tt];
            for (int s = 0; s < NS; s++) {
                int jump = L->specs[s].kind == 0 ? run_scalar(L, s, tt, model, saved, L->Temps[tt])
                         : L->specs[s].kind == 1 ? run_block(L, s, tt, model, saved, L->Temps[tt])
                         : L->specs[s].kind == 3 ? run_multinomial(L, s, tt, model, saved, L->Temps[tt])
                                                 : run_slice(L, s, tt, model, saved, L->Temps[tt]);
                if (L->recDec) L->recDec[((size_t)L->itRun * K + tt) * NS + s] = (unsigned char)jump;
            }
            orc_ctx c; c.m = m; c.mut = model->mut; c.lp = model->lp;
            L->logProbTemps[tt] = get_logprob_all(&c); /* APT:447 */
        }
        /* ---- swap phase (APT:452-484) */
        pswap_calc(L);
        for (int ii = 1; ii <= K; ii++) {
            int iFrom = (iter % 2 == 0) ? ii : K + 1 - ii;
            double ucat, uacc;
            draw_swap_u(L, ii - 1, &ucat, &uacc);
            int f = iFrom - 1;
            int t = rcat(&L->pSwap[f], K, K, ucat);
            double pProp = L->pSwap[f + t * K];
            double pRev = L->pSwap[t + f * K];
            double lMHR = (L->invTemps[t] - L->invTemps[f]) * (L->logProbTemps[f] - L->logProbTemps[t]) + log(pRev) - log(pProp);
            double lu = log(uacc);
            int acc = lu < lMHR;
            if (L->recSwap) L->recSwap[(size_t)L->itRun * K + (ii - 1)] = t | (acc << 16);
            if (acc) {
                orc_buf tb = L->rung[t]; L->rung[t] = L->rung[f]; L->rung[f] = tb;
                tb = L->saved[t]; L->saved[t] = L->saved[f]; L->saved[f] = tb;
                double tmp = L->logProbTemps[t]; L->logProbTemps[t] = L->logProbTemps[f]; L->logProbTemps[f] = tmp;
                for (int j = 0; j < K; j++) { /* rows */
                    tmp = L->pSwap[t + j * K]; L->pSwap[t + j * K] = L->pSwap[f + j * K]; L->pSwap[f + j * K] = tmp;
                }
                for (int j = 0; j < K; j++) { /* cols */
                    tmp = L->pSwap[j + t * K]; L->pSwap[j + t * K] = L->pSwap[j + f * K]; L->pSwap[j + f * K] = tmp;
                }
                L->accCountSwap[f + t * K] += 1;
            }
        }
        /* ---- temperature ladder adaptation (APT:489-505) */
        L->totalIters++;
        if (adaptTemps) {
            double gammaTSA = 1 / pow((double)L->totalIters / tuneTemper1 + 3, tuneTemper2);
            for (int i = 0; i < K; i++) L->TempsCurrent[i] = L->Temps[i];
            for (int ii = 0; ii < K - 1; ii++) {
                /* min(1, x) as R evaluates it and as nimble compiles it (pairmin: x1 < x2 ? x1 : x2) [NIMBLE-UNVERIFIED]: a
                 * NaN -- two neighbouring rungs both at logProb -Inf -- propagates into the ladder; C's fmin() would hide it */
                const double e_ = exp((L->invTemps[ii + 1] - L->invTemps[ii]) * (L->logProbTemps[ii] - L->logProbTemps[ii + 1]));
                double accProb = (1.0 < e_) ? 1.0 : e_;
                L->Temps[ii + 1] = L->Temps[ii] + (L->TempsCurrent[ii + 1] - L->TempsCurrent[ii]) * exp(gammaTSA * (accProb - 0.234));
                if (L->Temps[ii + 1] > L->ULT) L->Temps[ii + 1] = L->ULT + (ii + 1);
            }
            for (int ii = 0; ii < K - 1; ii++) L->invTemps[ii + 1] = 1 / L->Temps[ii + 1];
        }
        /* ---- output (APT:514-539) */
        if (iter % thin == 0) {
            long row = off + iter / thin - 1;
            write_row(L, &L->rung[0], L->mon, L->nmon, &L->samples[row * L->nmon]);
            memcpy(&L->tempTraj[(iter / thin - 1) * K], L->Temps, sizeof(double) * K);
            L->logProbs[row] = L->logProbTemps[0];
            if (L->monitorTmax) write_row(L, &L->rung[K - 1], L->mon, L->nmon, &L->samplesTmax[row * L->nmon]);
        }
        if (iter % thin2 == 0) {
            long row = off2 + iter / thin2 - 1;
            write_row(L, &L->rung[0], L->mon2, L->nmon2, &L->samples2[row * L->nmon2]);
            if (L->monitorTmax) write_row(L, &L->rung[K - 1], L->mon2, L->nmon2, &L->samples2Tmax[row * L->nmon2]);
        }
        L->rngIter++;
    }
    /* APT:548 */
    memcpy(L->model.mut, L->rung[0].mut, sizeof(double) * m->nmut);
    memcpy(L->model.lp, L->rung[0].lp, sizeof(double) * m->nlp);
    return 0;
}

/* run many independent ladders, one per thread (the reference itself is serial: APT:435-448) */
typedef struct {
    orc_ladder** Ls;
    int n, *next, rc;
    pthread_mutex_t* mu;
    long niter;
    int reset, resetMV, resetTempering, adaptTemps, thin, thin2;
    double t1, t2;
} orc_job;
static void* orc_worker(void* arg) {
    orc_job* j = (orc_job*)arg;
    for (;;) {
        pthread_mutex_lock(j->mu);
        int i = (*j->next)++;
        pthread_mutex_unlock(j->mu);
        if (i >= j->n) break;
        int r = orc_ladder_run(j->Ls[i], j->niter, j->reset, j->resetMV, j->resetTempering, j->adaptTemps, j->t1,
                               j->t2, j->thin, j->thin2);
        if (r) j->rc = r;
    }
    return NULL;
}
int orc_max_threads(void) {
    long n = sysconf(_SC_NPROCESSORS_ONLN);
    return n > 0 ? (int)n : 1;
}
int orc_ladders_run(orc_ladder** Ls, int n, int nthreads, long niter, int reset, int resetMV, int resetTempering,
                    int adaptTemps, double t1, double t2, int thin, int thin2) {
    if (nthreads <= 0) nthreads = orc_max_threads();
    if (nthreads > n) nthreads = n;
    if (nthreads < 1) nthreads = 1;
    pthread_mutex_t mu = PTHREAD_MUTEX_INITIALIZER;
    int next = 0;
    orc_job job = {Ls, n, &next, 0, &mu, niter, reset, resetMV, resetTempering, adaptTemps, thin, thin2, t1, t2};
    pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * nthreads);
    for (int t = 1; t < nthreads; t++) pthread_create(&th[t], NULL, orc_worker, &job);
    orc_worker(&job);
    for (int t = 1; t < nthreads; t++) pthread_join(th[t], NULL);
    free(th);
    return job.rc;
}

/* calculateWAIC  (APT:593-635).  data_lp[ndata]: logProb-arena indices of the data nodes (dataNodes, APT:323);
   nburnin_post = ceiling(nburnin / thin) is formed by the caller (APT:605).  Returns WAIC; out3 = {WAIC, logAvgProb,
   pWAIC}.  Fewer than 2 post-burn-in samples: -Inf (APT:607-610). */
double orc_ladder_waic(const orc_ladder* L, long nburnin_post, int ndata, const int* data_lp, double* out3) {
    const orc_model* m = L->m;
    const long n = L->nrows - nburnin_post; /* numMCMCSamples, APT:606 */
    if (n < 2) return -INFINITY;
    double* lpp = (double*)malloc(sizeof(double) * (size_t)n * ndata); /* logPredProbs, APT:611 */
    orc_ctx c;
    c.m = m;
    c.mut = (double*)malloc(sizeof(double) * (m->nmut + 1));
    c.lp = (double*)calloc(m->nlp + 1, sizeof(double));
    for (long i = 0; i < n; i++) {
        /* copy(mvSamples, model, nodesTo = sampledNodes, row = i + nburninPostThinning): unmonitored nodes keep the
           model's current values (APT:614,617); simulate(paramDeps) + calculate(dataNodes) (APT:618-619) */
        memcpy(c.mut, L->model.mut, sizeof(double) * m->nmut);
        const double* row = &L->samples[(size_t)(i + nburnin_post) * L->nmon];
        for (int q = 0; q < L->nmon; q++)
            if (L->mon[q] >= 0) c.mut[L->mon[q]] = row[q];
        calculate_all(&c);
        for (int j = 0; j < ndata; j++) lpp[(size_t)i * ndata + j] = c.lp[data_lp[j]];
    }
    double logAvgProb = 0, pWAIC = 0;
    for (int j = 0; j < ndata; j++) { /* APT:623-629 */
        double mx = -INFINITY;
        for (long i = 0; i < n; i++) { double v = lpp[(size_t)i * ndata + j]; if (v > mx || v != v) mx = v; }
        long double se = 0, sm = 0;
        for (long i = 0; i < n; i++) { se += exp(lpp[(size_t)i * ndata + j] - mx); sm += lpp[(size_t)i * ndata + j]; }
        logAvgProb += mx + log((double)(se / n));
        long double mean = sm / n, r = 0; /* R's mean(): a second pass refines the first */
        for (long i = 0; i < n; i++) r += lpp[(size_t)i * ndata + j] - mean;
        mean += r / n;
        long double sv = 0;
        for (long i = 0; i < n; i++) { long double d = lpp[(size_t)i * ndata + j] - mean; sv += d * d; }
        pWAIC += (double)(sv / (n - 1)); /* var() */
    }
    free(lpp); free(c.mut); free(c.lp);
    const double WAIC = -2 * (logAvgProb - pWAIC); /* APT:630 */
    if (out3) { out3[0] = WAIC; out3[1] = logAvgProb; out3[2] = pWAIC; }
    return WAIC;
}

/* ------------------------------------------------------------------ accessors */
long orc_ladder_nrows(const orc_ladder* L, int which) { return which == 2 ? L->nrows2 : which == 3 ? L->ttrows : L->nrows; }
const double* orc_ladder_ptr(const orc_ladder* L, int what) {
    switch (what) {
        case 0: return L->samples;
        case 1: return L->samples2;
        case 2: return L->samplesTmax;
        case 3: return L->samples2Tmax;
        case 4: return L->logProbs;
        case 5: return L->tempTraj;
        case 6: return L->Temps;
        case 7: return L->accCountSwap;
        case 8: return L->logProbTemps;
        default: return NULL;
    }
}
const double* orc_ladder_rung_state(const orc_ladder* L, int tt) { return L->rung[tt].mut; }
double orc_ladder_scale(const orc_ladder* L, int tt, int s) { return L->ss[s + tt * L->NS].scale; }
const double* orc_ladder_propcov(const orc_ladder* L, int tt, int s) { return L->ss[s + tt * L->NS].propCov; }
const double* orc_ladder_multinomial_state(const orc_ladder* L, int tt, int s) { return L->ss[s + tt * L->NS].mn; }
double orc_ladder_multinomial_u(const orc_ladder* L, int tt, int s) { return L->ss[s + tt * L->NS].u; }
long orc_ladder_total_iters(const orc_ladder* L) { return L->totalIters; }
