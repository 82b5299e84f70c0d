"""oracle/oracle.py — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

ctypes wrapper around oracle/liboracle.so (apt_oracle.c + rmath_port.c).  Imported only by tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` legs — never by the product
package `nimbleapt_b200`.
"""
import ctypes as C
import math
import os
import subprocess
import numpy as np

from .bugs_front import OracleModel

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE])


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(path):
            build()
        L = C.CDLL(path)
        dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int)
        L.orc_model_new.restype = C.c_void_p
        L.orc_model_new.argtypes = [C.c_int, ip, ip, C.c_int, C.c_int, dp, C.c_int, ip, C.c_int, ip, C.c_int, dp]
        L.orc_model_free.argtypes = [C.c_void_p]
        L.orc_model_logprob.restype = C.c_double
        L.orc_model_logprob.argtypes = [C.c_void_p, dp, dp]
        L.orc_model_logprob_ld.restype = C.c_double
        L.orc_model_logprob_ld.argtypes = [C.c_void_p, dp, dp]
        L.orc_specs_new.restype = C.c_void_p
        L.orc_specs_new.argtypes = [C.c_int]
        L.orc_spec_set.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, ip, C.c_int, ip, ip, C.c_int, ip,
                                   ip, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double,
                                   dp, C.c_int, C.c_int]
        L.orc_spec_set_slice.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
        L.orc_ladder_new.restype = C.c_void_p
        L.orc_ladder_new.argtypes = [C.c_void_p, C.c_void_p, C.c_int, dp, C.c_int, C.c_uint32, C.c_uint32, C.c_double,
                                     C.c_int, C.c_int, ip, C.c_int, ip]
        L.orc_ladder_free.argtypes = [C.c_void_p]
        L.orc_ladder_set_tables.argtypes = [C.c_void_p, dp, dp, dp, C.c_int, C.POINTER(C.c_ubyte), ip]
        L.orc_ladder_set_model.argtypes = [C.c_void_p, dp]
        L.orc_ladder_run.argtypes = [C.c_void_p, C.c_long, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double,
                                     C.c_int, C.c_int]
        L.orc_ladders_run.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_long, C.c_int, C.c_int, C.c_int,
                                      C.c_int, C.c_double, C.c_double, C.c_int, C.c_int]
        L.orc_max_threads.restype = C.c_int
        L.orc_ladder_nrows.restype = C.c_long
        L.orc_ladder_nrows.argtypes = [C.c_void_p, C.c_int]
        L.orc_ladder_ptr.restype = dp
        L.orc_ladder_ptr.argtypes = [C.c_void_p, C.c_int]
        L.orc_ladder_rung_state.restype = dp
        L.orc_ladder_rung_state.argtypes = [C.c_void_p, C.c_int]
        L.orc_ladder_scale.restype = C.c_double
        L.orc_ladder_scale.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.orc_ladder_propcov.restype = dp
        L.orc_ladder_propcov.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.orc_ladder_total_iters.restype = C.c_long
        L.orc_ladder_total_iters.argtypes = [C.c_void_p]
        L.orc_spec_set_multinomial.argtypes = [C.c_void_p, C.c_int, C.c_double]
        L.orc_ladder_multinomial_state.restype = dp
        L.orc_ladder_multinomial_state.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.orc_ladder_multinomial_u.restype = C.c_double
        L.orc_ladder_multinomial_u.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.orc_set_multinomial_exact_reverse.argtypes = [C.c_int]
        L.apt_dmulti_log.restype = C.c_double
        L.apt_dmulti_log.argtypes = [dp, dp, C.c_int, C.c_double]
        L.apt_rbinom_inv.restype = C.c_double
        L.apt_rbinom_inv.argtypes = [C.c_double] * 3
        L.orc_ladder_waic.restype = C.c_double
        L.orc_ladder_waic.argtypes = [C.c_void_p, C.c_long, C.c_int, ip, dp]
        for f in ("apt_dnorm_log", "apt_dunif_log", "apt_dgamma_log", "apt_dbeta_log", "apt_dbinom_log"):
            getattr(L, f).restype = C.c_double
            getattr(L, f).argtypes = [C.c_double] * 3
        for f in ("apt_dpois_log", "apt_dexp_log", "apt_bd0", "apt_lbeta"):
            getattr(L, f).restype = C.c_double
            getattr(L, f).argtypes = [C.c_double] * 2
        L.apt_stirlerr.restype = C.c_double
        L.apt_stirlerr.argtypes = [C.c_double]
        _LIB = L
    return _LIB


def _ia(a):
    a = np.ascontiguousarray(np.asarray(a, dtype=np.int32))
    return a, a.ctypes.data_as(C.POINTER(C.c_int))


def _da(a):
    a = np.ascontiguousarray(np.asarray(a, dtype=np.float64))
    return a, a.ctypes.data_as(C.POINTER(C.c_double))


SAMPLER_KINDS = {"sampler_RW_tempered": 0, "RW_tempered": 0, "sampler_RW_block_tempered": 1, "RW_block_tempered": 1,
                 "sampler_slice_tempered": 2, "slice_tempered": 2,
                 "sampler_RW_multinomial_tempered": 3, "RW_multinomial_tempered": 3}


class OracleAPT:
    """CPU oracle with the shape of the reference's buildAPT object (APT_functions.R:242-639)."""

    def __init__(self, code, constants=None, data=None, inits=None, samplers=(), monitors=(), monitors2=(),
                 Temps=(1.0,), ULT=1e6, monitorTmax=True, nLadders=1, seed=11, ladder0=0, functions=None):
        L = lib()
        self.om = om = OracleModel(code, constants, data, inits, functions)
        self.K = len(Temps)
        self.nLadders = nLadders
        specs = []
        for s in samplers:
            ctl = dict(s.get("control", {}))
            kind = SAMPLER_KINDS[s["type"]]
            if kind == 2:  # sampler_slice_tempered never reads temperPriors (APT_functions.R:909-912)
                ctl.setdefault("temperPriors", True)
            if kind == 3:  # sampler_RW_multinomial_tempered reads useTempering instead (APT_functions.R:1015)
                if "useTempering" not in ctl:
                    raise ValueError("control$useTempering unspecified in sampler_RW_multinomial_tempered")
                ctl["temperPriors"] = bool(ctl["useTempering"])
            if "temperPriors" not in ctl:  # APT_functions.R:660,781
                raise ValueError("control$temperPriors unspecified in %s" % s["type"])
            tgt = om.expand_target(s["target"])
            if kind == 2 and len(tgt) > 1:  # APT_functions.R:928
                raise ValueError("cannot use slice sampler on more than one target node")
            if kind == 0 and len(tgt) > 1:  # APT_functions.R:682
                raise ValueError("cannot use RW sampler on more than one target; try RW_block sampler")
            if kind == 3:
                if om.target_dist(tgt[0]) != "dmulti":  # APT_functions.R:1046
                    raise ValueError("can only use RW_multinomial sampler for multinomial distributions")
                if len(om.target_nodes(tgt)) > 1:  # APT_functions.R:1047
                    raise ValueError("cannot use RW_multinomial sampler on more than one target")
                if ctl.get("adaptive", True) and int(ctl.get("adaptInterval", 200)) < 100:  # APT_functions.R:1048
                    raise ValueError("adaptInterval < 100 is not recommended for RW_multinomial sampler")
            cd, ci, td, ti = om.dependencies(tgt)
            lower = upper = -1
            if kind == 0 and ctl.get("reflective", False):
                lo_t, hi_t, subst = om.target_bounds(tgt[0])
                lo_v = lo_t[1] if lo_t[0] == "num" else None
                hi_v = hi_t[1] if hi_t[0] == "num" else None
                if lo_v is None or np.isfinite(lo_v):
                    lower = om.compile_scalar(lo_t, [], subst)
                if hi_v is None or np.isfinite(hi_v):
                    upper = om.compile_scalar(hi_t, [], subst)
            specs.append(dict(kind=kind, tgt=tgt, cd=cd, ci=ci, td=td, ti=ti, ctl=ctl, lower=lower, upper=upper))
        om.finalize_code()
        keep = self._keep = []
        vo, vop = _ia([v.off for v in om.varlist]); vn, vnp = _ia([v.nrow for v in om.varlist])
        ar, arp = _da(om.arena0)
        dr, drp = _ia(np.asarray(om.decl_recs, dtype=np.int32).reshape(-1))
        cd_, cdp = _ia(om.code + [0, 0]); cs, csp = _da(om.consts + [0.0])
        self.model = L.orc_model_new(len(om.varlist), vop, vnp, om.nmut, om.ntot, arp, len(om.decl_recs), drp,
                                     len(cd_), cdp, len(cs), csp)
        self.NS = len(specs)
        self.specs = L.orc_specs_new(max(1, self.NS))
        self.dmax = 1
        for i, sp in enumerate(specs):
            ctl = sp["ctl"]
            d = len(sp["tgt"])
            self.dmax = max(self.dmax, d)
            t, tp = _ia(sp["tgt"]); a, ap = _ia(sp["cd"]); b, bp = _ia(sp["ci"]); c, cp = _ia(sp["td"]); e, ep = _ia(sp["ti"])
            pc = ctl.get("propCov", "identity")
            if isinstance(pc, str):
                pcp = None
            else:
                pcm = np.asarray(pc, dtype=float)
                if pcm.shape != (d, d):  # APT_functions.R:816
                    raise ValueError("propCov matrix must have dimension %d x %d" % (d, d))
                if not np.allclose(pcm, pcm.T):  # APT_functions.R:817
                    raise ValueError("propCov matrix must be symmetric")
                pca, pcp = _da(pcm.reshape(-1, order="F"))
            scale = float(ctl.get("scale", 1.0)) if sp["kind"] != 2 else float(ctl.get("width", 1.0))  # APT_functions.R:911
            afe = float(ctl.get("adaptFactorExponent", 0.8))
            if ctl.get("log", False) and ctl.get("reflective", False):  # APT_functions.R:684
                raise ValueError("cannot use reflective RW sampler on a log scale")
            if afe < 0 or scale < 0:  # APT_functions.R:685-686
                raise ValueError("adaptFactorExponent and scale must be non-negative")
            L.orc_spec_set(self.model, self.specs, i, sp["kind"], d, tp, len(a), ap, bp, len(c), cp, ep,
                           int(bool(ctl.get("log", False))), int(bool(ctl.get("reflective", False))),
                           int(bool(ctl.get("adaptive", True))), int(bool(ctl.get("adaptScaleOnly", False))),
                           int(ctl.get("adaptInterval", 200)), int(bool(ctl["temperPriors"])), afe, scale, pcp,
                           sp["lower"], sp["upper"])
            if sp["kind"] == 3:  # Ntotal <- sum(values(model, target)) at setup (APT_functions.R:1024)
                L.orc_spec_set_multinomial(self.specs, i, float(sum(om.arena0[t_] for t_ in sp["tgt"])))
            if sp["kind"] == 2:
                L.orc_spec_set_slice(self.specs, i, int(ctl.get("sliceMaxSteps", 100)), int(om.is_discrete(sp["tgt"][0])))
        self.monitors = list(monitors)
        self.mon, self.mon_labels = om.monitor_indices(list(monitors))
        self.mon2, self.mon2_labels = om.monitor_indices(list(monitors2))
        m1, m1p = _ia(self.mon if self.mon else [0]); m2, m2p = _ia(self.mon2 if self.mon2 else [0])
        tv, tvp = _da(Temps)
        self.ladders = [L.orc_ladder_new(self.model, self.specs, self.NS, tvp, self.K, seed, ladder0 + l, float(ULT),
                                         int(monitorTmax), len(self.mon), m1p, len(self.mon2), m2p)
                        for l in range(nLadders)]
        self._tables = None

    def __del__(self):
        try:
            L = lib()
            for h in self.ladders:
                L.orc_ladder_free(h)
            L.orc_model_free(self.model)
        except Exception:
            pass

    @property
    def nparam(self):
        return self.om.nparam

    def set_inits(self, theta, ladder=None):
        """theta: the PARAM prefix of the mutable arena (nparam doubles); deterministic values are recomputed."""
        L = lib()
        full = self.om.arena0[:self.om.nmut].copy()
        full[:len(theta)] = theta
        a, ap = _da(full)
        for l in (range(self.nLadders) if ladder is None else [ladder]):
            L.orc_ladder_set_model(self.ladders[l], ap)

    def set_tables(self, ladder, Z=None, U=None, S=None, record=False, niter=0):
        """Inject randoms (Z[niter,K,NS,dmax], U[niter,K,NS], S[niter,K,2]) and/or record decisions."""
        L = lib()
        keep = []
        def ptr(x):
            if x is None:
                return None
            a, p = _da(x); keep.append(a); return p
        rec_d = rec_s = None
        rdp = rsp = None
        if record:
            rec_d = np.zeros((niter, self.K, max(1, self.NS)), dtype=np.uint8)
            rec_s = np.zeros((niter, self.K), dtype=np.int32)
            rdp = rec_d.ctypes.data_as(C.POINTER(C.c_ubyte)); rsp = rec_s.ctypes.data_as(C.POINTER(C.c_int))
        L.orc_ladder_set_tables(self.ladders[ladder], ptr(Z), ptr(U), ptr(S), self.dmax, rdp, rsp)
        self._keep.append((keep, rec_d, rec_s))
        return rec_d, rec_s

    def run(self, niter, reset=True, resetMV=False, resetTempering=False, adaptTemps=True, tuneTemper1=10.0,
            tuneTemper2=1.0, thin=1, thin2=1, nthreads=1):
        L = lib()
        arr = (C.c_void_p * self.nLadders)(*self.ladders)
        rc = L.orc_ladders_run(arr, self.nLadders, nthreads, niter, int(reset), int(resetMV), int(resetTempering),
                               int(adaptTemps), float(tuneTemper1), float(tuneTemper2), int(thin), int(thin2))
        if rc:
            raise RuntimeError("cannot specify niter < 0")

    def _arr(self, ladder, what, rows, cols):
        L = lib()
        p = L.orc_ladder_ptr(self.ladders[ladder], what)
        n = rows * cols
        if n == 0:
            return np.zeros((rows, cols))
        return np.ctypeslib.as_array(p, shape=(n,)).copy().reshape(rows, cols)

    def nrows(self, ladder=0, which=1):
        return lib().orc_ladder_nrows(self.ladders[ladder], which)

    def samples(self, ladder=0):
        return self._arr(ladder, 0, self.nrows(ladder, 1), len(self.mon))

    def samples2(self, ladder=0):
        return self._arr(ladder, 1, self.nrows(ladder, 2), len(self.mon2))

    def samplesTmax(self, ladder=0):
        return self._arr(ladder, 2, self.nrows(ladder, 1), len(self.mon))

    def logProbs(self, ladder=0):
        return self._arr(ladder, 4, self.nrows(ladder, 1), 1)[:, 0]

    def tempTraj(self, ladder=0):
        return self._arr(ladder, 5, self.nrows(ladder, 3), self.K)

    def Temps(self, ladder=0):
        return self._arr(ladder, 6, 1, self.K)[0]

    def accCountSwap(self, ladder=0):
        return self._arr(ladder, 7, self.K, self.K).T  # stored column-major [from + to*K]

    def logProbTemps(self, ladder=0):
        return self._arr(ladder, 8, 1, self.K)[0]

    def rung_state(self, ladder, tt):
        p = lib().orc_ladder_rung_state(self.ladders[ladder], tt)
        return np.ctypeslib.as_array(p, shape=(self.om.nmut,)).copy()[:self.om.nparam]

    def scale(self, ladder, tt, s):
        return lib().orc_ladder_scale(self.ladders[ladder], tt, s)

    def multinomialState(self, ladder, tt, s, d):
        """[u, timesRan, AcceptRates, ScaleShifts, totalAdapted, timesAccepted, ENSwapMatrix, ENSwapDeltaMatrix,
        RescaleThreshold] of a sampler_RW_multinomial_tempered (APT_functions.R:1029-1040); matrices column-major."""
        p = lib().orc_ladder_multinomial_state(self.ladders[ladder], tt, s)
        m = np.ctypeslib.as_array(p, shape=(8 * d * d,)).copy()
        return np.concatenate([[lib().orc_ladder_multinomial_u(self.ladders[ladder], tt, s)], m])

    def propCov(self, ladder, tt, s, d):
        p = lib().orc_ladder_propcov(self.ladders[ladder], tt, s)
        return np.ctypeslib.as_array(p, shape=(d * d,)).copy().reshape(d, d, order="F")

    def calculateWAIC(self, nburnin=0, ladder=0, thin=1):
        """calculateWAIC (APT_functions.R:593-635) on the rung-1 samples of one ladder; `thin` is the thinning
        interval of the run that produced them (thinToUseVec[1], APT_functions.R:605).  Returns (WAIC, lppd, pWAIC)."""
        idx = self.om.data_node_lp_indices()
        if not idx:  # APT_functions.R:331
            raise ValueError("WAIC cannot be calculated, as no data nodes were detected in the model.")
        self.om.check_waic_monitors(self.monitors)
        a, ap = _ia(idx)
        out = np.zeros(3)
        w = lib().orc_ladder_waic(self.ladders[ladder], int(math.ceil(nburnin / thin)), len(idx), ap,
                                  out.ctypes.data_as(C.POINTER(C.c_double)))
        return (w, out[1], out[2]) if np.isfinite(w) or w != w else (w, float("nan"), float("nan"))

    def logprob(self, theta, extended=False):
        """Whole-model log probability at PARAM vector theta; returns (total, per-declaration sums).  extended=True adds
        the same node terms in long double instead of sequentially in double like model$getLogProb()."""
        full = self.om.arena0[:self.om.nmut].copy()
        full[:len(theta)] = theta
        a, ap = _da(full)
        out = np.zeros(len(self.om.decl_recs))
        fn = lib().orc_model_logprob_ld if extended else lib().orc_model_logprob
        tot = fn(self.model, ap, out.ctypes.data_as(C.POINTER(C.c_double)))
        return tot, out
