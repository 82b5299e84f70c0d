"""Host-side mirror of the reference's user interface, on top of the C ABI (include/aptb200.h).

The reference's host language is R (absent from this image), so this module restates the R surface in Python
with the same names, argument meanings and error behaviour, so that tests read like the reference's own usage
(/root/reference/nimbleAPT/vignettes/nimbleAPT-vignette.Rmd:127-141, demo/APT_demo.R:170-219):

    code  = nimbleCode('''{ ... BUGS ... }''')
    model = nimbleModel(code, constants=..., data=..., inits=...)
    conf  = configureMCMC(model, nodes="centroids", monitors="centroids")
    conf.removeSamplers()
    conf.addSampler("centroids[1]", type="sampler_RW_tempered", control=dict(temperPriors=True))
    apt   = buildAPT(conf, Temps=range(1, 6), ULT=1000)
    capt  = compileNimble(apt)              # no-op: buildAPT already lowered + compiled with nvcc
    capt.run(niter=15000)                   # cAPT$run(...)  APT_functions.R:336-348
    samples = as_matrix(capt.mvSamples)     # rows = thinned iterations, cols = expanded monitor names

`runAPT(cAPT, niter, thin, ...)` named by the north star does not exist in the reference; it is
provided as a thin wrapper over `run`.

Only device memory, kernels and NCCL live below this file, all inside libaptb200.so / the lowered modules.
numpy is used for host arrays; torch is not used.
"""
import ctypes as C
import os
import numpy as np

from . import _lib
from ._lib import AptError, BuildOpts, RunArgs, SamplerControl, Timing, check, lib

_SAMPLER_TYPES = ("sampler_RW_tempered", "RW_tempered", "sampler_RW_block_tempered", "RW_block_tempered",
                  "sampler_slice_tempered", "slice_tempered",
                  "sampler_RW_multinomial_tempered", "RW_multinomial_tempered")
_SCALAR_KEYS = {"log", "reflective", "adaptive", "adaptInterval", "adaptFactorExponent", "scale", "temperPriors"}
_BLOCK_KEYS = {"adaptive", "adaptScaleOnly", "adaptInterval", "adaptFactorExponent", "scale", "propCov", "temperPriors"}
# APT:909-912 (the slice sampler never reads temperPriors; the reference's demo passes it anyway, demo/APT_demo.R:176)
_SLICE_KEYS = {"adaptive", "adaptInterval", "width", "sliceMaxSteps", "temperPriors"}
_MULTI_KEYS = {"adaptive", "adaptInterval", "useTempering"}  # APT:1013-1015

(APT_SAMPLES, APT_SAMPLES2, APT_SAMPLES_TMAX, APT_SAMPLES2_TMAX, APT_LOGPROBS, APT_TEMPTRAJ, APT_TEMPS, APT_ACCSWAP,
 APT_LOGPROBTEMPS, APT_RUNG_STATES, APT_SCALES, APT_BLOCK_STATE, APT_DECISIONS, APT_SWAP_RECORD, APT_MODEL_STATE) = range(15)
(INFO_N_PARAMS, INFO_N_TEMPS, INFO_N_UNITS, INFO_N_GROUPS, INFO_N_MON, INFO_N_MON2, INFO_DMAX, INFO_ROWS, INFO_ROWS2,
 INFO_TTROWS, INFO_L, INFO_LAUNCHES, INFO_TOTAL_ITERS, INFO_W, INFO_CTA, INFO_ENGINE, INFO_BLKSZ, INFO_FP64_GFLOPS) = range(18)


def nimbleCode(text):
    """nimbleCode({...}): the BUGS text is kept verbatim and parsed by the library."""
    return str(text)


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


_FUNCTIONS = {}  # user-supplied functions by name: R finds them in the calling environment, here they register themselves


class NimbleFunction:
    """nimbleFunction(run = function(...) {...}): a deterministic function the model code calls (demo/APT_demo.R:42-71).
    `run` is the R text of the function; the library inlines it into the lowered model (apt_model_add_function)."""

    def __init__(self, run, name=None):
        self.run, self.name = str(run), name
        if name:
            _FUNCTIONS[name] = self


def nimbleFunction(run, name=None, **_ignored):
    return NimbleFunction(run, name)


class NimbleModel:
    """nimbleModel(code, constants, data, inits) for the supported BUGS subset."""

    def __init__(self, code, constants=None, data=None, inits=None, functions=None):
        self.code = code
        self.constants = {k: np.asarray(v, dtype=np.float64) for k, v in (constants or {}).items()}
        self.data = {k: np.asarray(v, dtype=np.float64) for k, v in (data or {}).items()}
        self.inits = {k: np.asarray(v, dtype=np.float64) for k, v in (inits or {}).items()}
        fns = {k: v for k, v in _FUNCTIONS.items() if (k + "(") in code.replace(" ", "")}
        fns.update(functions or {})
        self.functions = {k: (v.run if isinstance(v, NimbleFunction) else str(v)) for k, v in fns.items()}
        # parse now so that syntax errors surface where nimbleModel() would raise them
        h = C.c_void_p()
        check(lib().apt_model_create(code.encode(), C.byref(h)))
        try:
            for k, v in self.functions.items():
                check(lib().apt_model_add_function(h, k.encode(), v.encode()))
        finally:
            lib().apt_model_free(h)


def nimbleModel(code, constants=None, data=None, inits=None, functions=None, **_ignored):
    return NimbleModel(code, constants, data, inits, functions)


class MCMCconf:
    """configureMCMC(model, ...): holds samplerConfs, monitors, thin (APT_functions.R:249-250, 302-311)."""

    def __init__(self, model, nodes=None, monitors=None, monitors2=None, thin=1, thin2=1, enableWAIC=False, **_ignored):
        if not isinstance(model, NimbleModel):
            raise TypeError("configureMCMC needs a nimbleModel")
        self.model = model
        self.monitors = [monitors] if isinstance(monitors, str) else list(monitors or [])
        self.monitors2 = [monitors2] if isinstance(monitors2, str) else list(monitors2 or [])
        self.thin, self.thin2 = int(thin), int(thin2)
        self.enableWAIC = bool(enableWAIC)  # APT:231,329
        # nimble would assign default (non-tempered) samplers here; buildAPT can only use tempered samplers, so
        # like the vignette (:130) users call removeSamplers() and add tempered ones.
        self.samplerConfs = []

    def removeSamplers(self, *a, **k):
        self.samplerConfs = []

    def addSampler(self, target, type="sampler_RW_tempered", control=None, **_ignored):
        control = dict(control or {})
        name = getattr(type, "__name__", type)
        if name not in _SAMPLER_TYPES:
            raise ValueError("unsupported sampler type %r (supported: sampler_RW_tempered, sampler_RW_block_tempered, "
                             "sampler_slice_tempered, sampler_RW_multinomial_tempered)" % (name,))
        block = "block" in name
        unknown = set(control) - (_MULTI_KEYS if "multinomial" in name else _SLICE_KEYS if "slice" in name else _BLOCK_KEYS if block else _SCALAR_KEYS)
        if unknown:
            raise ValueError("unknown control option(s) for %s: %s" % (name, ", ".join(sorted(unknown))))
        if isinstance(target, (list, tuple)):
            target = "c(" + ",".join("'%s'" % t for t in target) + ")"
        self.samplerConfs.append(dict(target=target, type=name, control=control))

    def addMonitors(self, *names):
        self.monitors += [n for n in names]

    def addMonitors2(self, *names):
        self.monitors2 += [n for n in names]

    def printSamplers(self):
        for i, s in enumerate(self.samplerConfs):
            print("[%d] %s sampler: %s" % (i + 1, s["type"].replace("sampler_", ""), s["target"]))


def configureMCMC(model, **kw):
    return MCMCconf(model, **kw)


# the two exported sampler generators of the reference (nimbleAPT/NAMESPACE:6,8) are selected by name
def sampler_RW_tempered():  # pragma: no cover - marker object
    raise AptError("sampler_RW_tempered is selected by name in conf.addSampler(type=...)")


def sampler_RW_block_tempered():  # pragma: no cover - marker object
    raise AptError("sampler_RW_block_tempered is selected by name in conf.addSampler(type=...)")


def sampler_RW_multinomial_tempered():  # pragma: no cover - marker object
    raise AptError("sampler_RW_multinomial_tempered is selected by name in conf.addSampler(type=...)")


def sampler_slice_tempered():  # pragma: no cover - marker object
    raise AptError("sampler_slice_tempered is selected by name in conf.addSampler(type=...)")


class _Samples:
    """Stand-in for a nimble modelValues object holding MCMC output; `as_matrix(x)` mirrors as.matrix()."""

    def __init__(self, apt, what, names):
        self._apt, self._what, self.names = apt, what, names

    def matrix_column_major(self, ladder=0):
        """The same matrix of one ladder delivered in column-major order (what an R matrix is): transposed on the device,
        copied as it is (apt_get_layout with APT_COLUMN_MAJOR).  Returned as a Fortran-ordered [rows, monitors] array."""
        rows = self._apt.info(INFO_ROWS2 if self._what in (APT_SAMPLES2, APT_SAMPLES2_TMAX) else INFO_ROWS)
        cols = len(self.names)
        out = np.empty(max(rows * cols, 1), dtype=np.float64)
        got = C.c_int64()
        check(lib().apt_get_layout(self._apt._h, self._what, int(ladder), 1, _dp(out), out.size, C.byref(got)))
        return out[:got.value].reshape((rows, cols), order="F")

    def matrix(self, ladder=0, out=None):
        """as.matrix(cAPT$mvSamples) of one ladder ([rows, monitors]) or of all local ladders (ladder = -1:
        [ladders, rows, monitors]).  `out`: optional caller-allocated flat float64 destination (the C ABI always
        writes into caller-allocated memory; page-locked memory makes the device-to-host copy several times faster)."""
        return self._apt._get_trace(self._what, ladder, len(self.names), out)

    def __array__(self, dtype=None, copy=None):
        return self.matrix(0)

    def __len__(self):
        return self.matrix(0).shape[0]


def as_matrix(x, ladder=0):
    return x.matrix(ladder) if isinstance(x, _Samples) else np.asarray(x)


class APT:
    """buildAPT(conf, Temps, monitorTmax, ULT, thinPrintTemps, ...)  (APT_functions.R:242-334) — the lowering,
    the nvcc compile and the device allocation all happen here, so `compileNimble` is an identity."""

    def __init__(self, conf, Temps, monitorTmax=True, ULT=1e6, thinPrintTemps=1, nLadders=1, seed=11, ladder0=0,
                 device=0, record=False, fuseLinks=True, lanesPerChain=0, ctaThreads=0, engine=0, rungsPerPass=0, clusterCtas=0,
                 verbose=False, compileOnly=False, enableWAIC=False, **dots):
        if isinstance(conf, NimbleModel):  # APT:249-250
            conf = configureMCMC(conf, **dots)
        elif not isinstance(conf, MCMCconf):  # APT:251-253
            raise TypeError("conf must either be a nimbleModel or a MCMCconf object (created by configureMCMC(...) )")
        L = lib()
        self._h = C.c_void_p()
        self._m = C.c_void_p()
        self.conf = conf
        model = conf.model
        check(L.apt_model_create(model.code.encode(), C.byref(self._m)))
        for k, v in model.functions.items():
            check(L.apt_model_add_function(self._m, k.encode(), v.encode()))
        for role, d in ((0, model.constants), (1, model.data), (2, model.inits)):
            for name, arr in d.items():
                a = np.asfortranarray(arr, dtype=np.float64)
                dims = (C.c_int64 * max(a.ndim, 1))(*(a.shape if a.ndim else (1,)))
                flat = np.ascontiguousarray(a.reshape(-1, order="F"))
                check(L.apt_model_set_array(self._m, name.encode(), role, _dp(flat), dims, a.ndim))
        self._keep = []
        for s in conf.samplerConfs:
            ctl = SamplerControl()
            L.apt_sampler_control_default(C.byref(ctl))
            c = s["control"]
            for k in ("log", "reflective", "adaptive", "adaptScaleOnly", "adaptInterval", "temperPriors", "sliceMaxSteps", "useTempering"):
                if k in c and c[k] is not None:
                    setattr(ctl, k, int(c[k]))
            for k in ("adaptFactorExponent", "scale", "width"):
                if k in c and c[k] is not None:
                    setattr(ctl, k, float(c[k]))
            pc = c.get("propCov", "identity")
            if not (isinstance(pc, str) and pc == "identity"):
                pcm = np.asarray(pc)
                if pcm.ndim != 2:
                    raise AptError("propCov must be a matrix\n")  # APT:814
                if not np.issubdtype(pcm.dtype, np.number):
                    raise AptError("propCov matrix must be numeric\n")  # APT:815
                if pcm.shape[0] != pcm.shape[1]:
                    raise AptError("propCov matrix must have dimension %dx%d\n" % (pcm.shape[0], pcm.shape[0]))
                flat = np.ascontiguousarray(pcm.astype(np.float64).reshape(-1, order="F"))
                self._keep.append(flat)
                ctl.propCov = _dp(flat)
                ctl.propCov_dim = pcm.shape[0]
            check(L.apt_model_add_sampler(self._m, s["target"].encode(), s["type"].encode(), C.byref(ctl)))
        for which, mons in ((1, conf.monitors), (2, conf.monitors2)):
            if mons:
                arr = (C.c_char_p * len(mons))(*[m.encode() for m in mons])
                check(L.apt_model_set_monitors(self._m, which, arr, len(mons)))
        check(L.apt_model_set_thin(self._m, conf.thin, conf.thin2))
        temps = np.ascontiguousarray(np.asarray(list(Temps), dtype=np.float64))
        self.opts = o = BuildOpts()
        L.apt_build_opts_default(C.byref(o))
        o.monitorTmax, o.ULT, o.n_ladders, o.ladder0, o.seed, o.device = int(bool(monitorTmax)), float(ULT), int(nLadders), int(ladder0), int(seed), int(device)
        o.record_decisions, o.fuse_links, o.lanes_per_chain, o.cta_threads = int(bool(record)), int(bool(fuseLinks)), int(lanesPerChain), int(ctaThreads)
        o.engine, o.rungs_per_pass, o.verbose = int(engine), int(rungsPerPass), int(bool(verbose))
        o.cluster_ctas = int(clusterCtas)
        self.enableWAIC = bool(enableWAIC) or conf.enableWAIC  # APT:259,329
        o.enable_waic = int(self.enableWAIC)
        self._temps0 = temps
        self.nTemps = len(temps)
        self.nLadders = int(nLadders)
        self.thinPrintTemps = thinPrintTemps
        self.monitorTmax = bool(monitorTmax)
        self.paramNames = self._names(0)
        self.monitorNames = self._names(1)
        self.monitorNames2 = self._names(2)
        if compileOnly:  # lower + nvcc only (works without a GPU): used by __graft_entry__.build()
            buf = C.create_string_buffer(4096)
            check(L.apt_model_compile(self._m, len(temps), C.byref(o), buf, len(buf)))
            self.modulePath = buf.value.decode()
            return
        if os.environ.get("APTB200_PREBUILD"):  # build tool (tools/prebuild_test_modules.sh): fill the module cache, never run
            buf = C.create_string_buffer(4096)
            check(L.apt_model_compile(self._m, len(temps), C.byref(o), buf, len(buf)))
            raise AptError("APTB200_PREBUILD is set: the module was compiled into the cache and is not run")
        check(L.apt_build(self._m, _dp(temps), len(temps), C.byref(o), C.byref(self._h)))
        self.mvSamples = _Samples(self, APT_SAMPLES, self.monitorNames)
        self.mvSamples2 = _Samples(self, APT_SAMPLES2, self.monitorNames2)
        if monitorTmax:
            self.mvSamplesTmax = _Samples(self, APT_SAMPLES_TMAX, self.monitorNames)
            self.mvSamples2Tmax = _Samples(self, APT_SAMPLES2_TMAX, self.monitorNames2)

    # ---- helpers
    def _names(self, which):
        buf = C.create_string_buffer(1 << 20)
        check(lib().apt_model_names(self._m, which, buf, len(buf)))
        s = buf.value.decode()
        return s.split("\n") if s else []

    def lowered_source(self):
        """The CUDA source the lowering step emitted for this model."""
        need = C.c_int64()
        check(lib().apt_model_lower(self._m, self.nTemps, C.byref(self.opts), None, 0, C.byref(need)))
        buf = C.create_string_buffer(need.value)
        check(lib().apt_model_lower(self._m, self.nTemps, C.byref(self.opts), buf, need.value, C.byref(need)))
        return buf.value.decode()

    def info(self, what):
        return int(lib().apt_info(self._h, what))

    def _get(self, what, ladder, n, out=None):
        if out is None:
            out = np.empty(max(int(n), 1), dtype=np.float64)
        else:  # caller-allocated destination (e.g. page-locked memory): flat float64, at least n elements
            if out.dtype != np.float64 or not out.flags.c_contiguous or out.ndim != 1 or out.size < n:
                raise ValueError("out must be a flat, contiguous float64 array of at least %d elements" % n)
        got = C.c_int64()
        check(lib().apt_get(self._h, what, ladder, _dp(out), int(n), C.byref(got)))
        return out[:got.value]

    def _get_trace(self, what, ladder, cols, out=None):
        rows = self.info(INFO_ROWS2 if what in (APT_SAMPLES2, APT_SAMPLES2_TMAX) else INFO_ROWS)
        nl = self.nLadders if ladder < 0 else 1
        a = self._get(what, ladder, nl * rows * cols, out)
        return a.reshape((nl, rows, cols) if ladder < 0 else (rows, cols))

    def __del__(self):
        try:
            if self._h:
                lib().apt_free(self._h)
            if self._m:
                lib().apt_model_free(self._m)
        except Exception:
            pass

    # ---- cAPT$run  (APT_functions.R:336-549)
    def run(self, niter, reset=True, resetMV=False, resetTempering=False, simulateAll=False, time=False,
            adaptTemps=True, printTemps=False, tuneTemper1=10, tuneTemper2=1, progressBar=True, thin=-1, thin2=-1,
            itersPerLaunch=0, samplesOut=None, gatherRoot=None, _inject=None):
        ra = RunArgs()
        lib().apt_run_args_default(C.byref(ra))
        ra.niter, ra.reset, ra.resetMV, ra.resetTempering = int(niter), int(bool(reset)), int(bool(resetMV)), int(bool(resetTempering))
        ra.simulateAll, ra.time, ra.adaptTemps, ra.printTemps = int(bool(simulateAll)), int(bool(time)), int(bool(adaptTemps)), int(bool(printTemps))
        ra.tuneTemper1, ra.tuneTemper2, ra.progressBar, ra.thin, ra.thin2 = float(tuneTemper1), float(tuneTemper2), int(bool(progressBar)), int(thin), int(thin2)
        ra.iters_per_launch = int(itersPerLaunch)
        keep = []
        if _inject:
            for key, field in (("Z", "inject_normals"), ("U", "inject_accept_u"), ("S", "inject_swap_u")):
                if _inject.get(key) is not None:
                    a = np.ascontiguousarray(_inject[key], dtype=np.float64)
                    keep.append(a)
                    setattr(ra, field, _dp(a))
        if samplesOut is not None:
            # asynchronous copy-back (apt_run_args.samples_out): this run's rows of mvSamples, [ladders][niter / thin][monitors],
            # arrive in the caller's flat float64 buffer piece by piece while the run is still sampling
            if samplesOut.dtype != np.float64 or not samplesOut.flags.c_contiguous or samplesOut.ndim != 1:
                raise ValueError("samplesOut must be a flat, contiguous float64 array")
            ra.samples_out = _dp(samplesOut)
            ra.samples_out_capacity = samplesOut.size
        if gatherRoot is not None:
            # multi-GPU (after commInit): this run's rows of mvSamples travel to rank `gatherRoot` over NCCL piece by piece
            # while the run goes on; on the root they land in samplesOut, [ladders of rank 0 | rank 1 | ...][niter / thin][monitors]
            ra.gather, ra.gather_root = 1, int(gatherRoot)
        check(lib().apt_run(self._h, C.byref(ra)))

    # ---- multi-GPU: one process per GPU, ladders sharded by global id; NCCL only gathers traces / pools statistics
    @staticmethod
    def commUniqueId():
        """The 128-byte ncclUniqueId rank 0 creates and hands to the other ranks out of band."""
        uid = (C.c_ubyte * 128)()
        check(lib().apt_comm_unique_id(uid))
        return bytes(uid)

    def commInit(self, world_size, rank, unique_id):
        check(lib().apt_comm_init(self._h, int(world_size), int(rank), (C.c_ubyte * 128)(*unique_id)))
        self._comm = (int(world_size), int(rank))

    def commAllReduceSum(self, vec):
        buf = np.ascontiguousarray(vec, dtype=np.float64).copy()
        check(lib().apt_comm_allreduce_sum(self._h, _dp(buf), buf.size))
        return buf

    def commFinalize(self):
        check(lib().apt_comm_finalize(self._h))

    def getTimes(self):
        """getTimes() (APT:552-555) reports per-sampler wall times in the reference; here: CUDA-event timing."""
        t = Timing()
        check(lib().apt_get_timing(self._h, C.byref(t)))
        return dict(kernel_ms=t.kernel_ms, main_kernel_ms=t.main_kernel_ms, launches=t.launches, main_launches=t.main_launches,
                    run_wall_ms=t.run_wall_ms, gather_tail_ms=t.gather_tail_ms)

    # ---- member data read after run
    @property
    def logProbs(self):
        return self._get(APT_LOGPROBS, 0, self.info(INFO_ROWS))

    def logProbsLadder(self, ladder):
        return self._get(APT_LOGPROBS, ladder, self.info(INFO_ROWS))

    @property
    def tempTraj(self):
        return self.tempTrajLadder(0)

    def tempTrajLadder(self, ladder):
        r = self.info(INFO_TTROWS)
        return self._get(APT_TEMPTRAJ, ladder, r * self.nTemps).reshape(r, self.nTemps)

    @property
    def Temps(self):
        return self._get(APT_TEMPS, 0, self.nTemps)

    def TempsLadder(self, ladder):
        return self._get(APT_TEMPS, ladder, self.nTemps)

    @property
    def accCountSwap(self):
        return self.accCountSwapLadder(0)

    def accCountSwapLadder(self, ladder):
        return self._get(APT_ACCSWAP, ladder, self.nTemps ** 2).reshape(self.nTemps, self.nTemps)

    @property
    def logProbTemps(self):
        return self._get(APT_LOGPROBTEMPS, 0, self.nTemps)

    def rungStates(self, ladder=0):
        p = self.info(INFO_N_PARAMS)
        return self._get(APT_RUNG_STATES, ladder, self.nTemps * p).reshape(self.nTemps, p)

    def scales(self, ladder=0):
        nu = self.info(INFO_N_UNITS)
        return self._get(APT_SCALES, ladder, self.nTemps * nu).reshape(self.nTemps, nu)

    def blockState(self, ladder=0):
        b = self.info(INFO_BLKSZ)
        return self._get(APT_BLOCK_STATE, ladder, self.nTemps * b).reshape(self.nTemps, b)

    def decisions(self, ladder=0, niter=None):
        nu = self.info(INFO_N_UNITS)
        a = self._get(APT_DECISIONS, ladder, niter * self.nTemps * nu)
        return a.reshape(-1, self.nTemps, nu).astype(np.uint8)

    def swapRecord(self, ladder=0, niter=None):
        a = self._get(APT_SWAP_RECORD, ladder, niter * self.nTemps)
        return a.reshape(-1, self.nTemps).astype(np.int32)

    def modelState(self, ladder=0):
        return self._get(APT_MODEL_STATE, ladder, self.info(INFO_N_PARAMS))

    def calculate(self, theta):
        """model$calculate() at given parameter vectors: returns (total logProb [n], per-declaration sums [n, ND])."""
        th = np.ascontiguousarray(np.atleast_2d(theta), dtype=np.float64)
        n, nd = th.shape[0], self.info(INFO_N_GROUPS)
        tot, grp = np.empty(n), np.empty((n, nd))
        check(lib().apt_logprob(self._h, _dp(th), n, _dp(tot), _dp(grp)))
        return tot, grp

    def calculateWAIC(self, nburnin=0, ladder=0, details=False):
        """cAPT$calculateWAIC(nburnin) (APT_functions.R:593-635), evaluated on the device from mvSamples.
        ladder = -1 pools the samples of all local ladders.  details=True returns (WAIC, lppd, pWAIC)."""
        if not self.enableWAIC:  # APT:596-599: prints and returns NaN
            print("Error: must set enableWAIC = TRUE in buildAPT to use the method calcualteWAIC. See help(buildAPT) and "
                  "help(buildMCMC) for additional information.")
            return float("nan")
        out = np.empty(3)
        check(lib().apt_waic(self._h, int(nburnin), int(ladder), _dp(out)))
        if out[0] == -np.inf:  # APT:607-610
            print("Error: need more than one post burn-in MCMC samples")
        elif np.isnan(out[0]):  # APT:632
            print("WAIC was calculated as NaN.  You may need to add monitors to model latent states, in order for a valid "
                  "WAIC calculation.")
        return tuple(out) if details else float(out[0])

    def setInits(self, theta, ladder=-1):
        th = np.ascontiguousarray(theta, dtype=np.float64)
        check(lib().apt_set_inits(self._h, _dp(th), ladder))

    def setData(self, name, values):
        a = np.ascontiguousarray(np.asfortranarray(values, dtype=np.float64).reshape(-1, order="F"))
        check(lib().apt_set_array(self._h, name.encode(), _dp(a), a.size))


def buildAPT(conf, Temps, monitorTmax=True, ULT=1e6, thinPrintTemps=1, **kw):
    return APT(conf, Temps, monitorTmax=monitorTmax, ULT=ULT, thinPrintTemps=thinPrintTemps, **kw)


def compileNimble(x, *more, **kw):
    """compileNimble(aptR): identity — the model was lowered and compiled with nvcc inside buildAPT."""
    return x if not more else (x,) + more


def runAPT(cAPT, niter, thin=-1, **kw):
    """runAPT(cAPT, niter, thin, ...): convenience wrapper over cAPT$run (not present in the reference, F5)."""
    cAPT.run(niter, thin=thin, **kw)
    return as_matrix(cAPT.mvSamples)


def plotTempTraj(cAPT):  # APT:1290-1308 — plotting is out of scope; return what the plot would show
    return cAPT.tempTraj
