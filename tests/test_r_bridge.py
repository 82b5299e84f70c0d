"""The .Call bridge (nimbleapt_b200/R/bridge.c) compiled and executed without R.

This image has no R, so the bridge is built against tests/rstub/ — a stand-in for the few entry points of
R's C API it uses (SEXP vectors with names/dim attributes, external pointers with finalizers, Rf_error unwinding by
longjmp, R_ToplevelExec / R_CheckUserInterrupt) — and driven from Python the way aptb200.R drives it through .Call.
What this checks is the marshalling: argument conversion, control-list keys, matrices arriving in R's column-major layout
(written by the library itself, apt_get_layout), caller-allocated outputs, samplesOut / setInits / setData / calculate, that a non-zero status becomes an R error only AFTER the library call returned with the
protection stack balanced, and that finalizers release the handles.  It does not check aptb200.R itself."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import models

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STUB = os.path.join(ROOT, "tests", "rstub")
SO = os.path.join(STUB, "_rbridge_test.so")
SEXP = C.c_void_p


def _build():
    srcs = [os.path.join(ROOT, "nimbleapt_b200", "R", "bridge.c"), os.path.join(STUB, "rstub.c")]
    deps = srcs + [os.path.join(STUB, "Rinternals.h"), os.path.join(ROOT, "include", "aptb200.h")]
    if not os.path.exists(os.path.join(ROOT, "nimbleapt_b200", "libaptb200.so")):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "nimbleapt_b200", "csrc")])
    if not os.path.exists(SO) or any(os.path.getmtime(d) > os.path.getmtime(SO) for d in deps):
        subprocess.check_call(["gcc", "-O1", "-Wall", "-Werror", "-Wno-strict-prototypes", "-fPIC", "-shared", "-I" + STUB,
                               "-I" + os.path.join(ROOT, "include"), "-o", SO] + srcs +
                              ["-L" + os.path.join(ROOT, "nimbleapt_b200"), "-laptb200",
                               "-Wl,-rpath," + os.path.join(ROOT, "nimbleapt_b200")])
    return C.CDLL(SO)


class R:
    """The handful of things the R interpreter does for aptb200.R: build argument objects, .Call, read results."""

    def __init__(self):
        L = self.L = _build()
        for f in ("stub_real", "stub_int", "stub_lgl", "stub_str", "stub_list", "stub_nil", "stub_call", "stub_data"):
            getattr(L, f).restype = SEXP
        L.stub_real.argtypes = [C.POINTER(C.c_double), C.c_ssize_t]
        L.stub_int.argtypes = L.stub_lgl.argtypes = [C.POINTER(C.c_int), C.c_ssize_t]
        L.stub_str.argtypes = [C.POINTER(C.c_char_p), C.c_ssize_t]
        L.stub_list.argtypes = [C.POINTER(C.c_char_p), C.POINTER(SEXP), C.c_ssize_t]
        L.stub_set_dim.argtypes = [SEXP, C.POINTER(C.c_int), C.c_int]
        L.stub_call.argtypes = [C.c_void_p, C.c_int, C.POINTER(SEXP)]
        L.stub_type.argtypes = L.stub_data.argtypes = [SEXP]
        L.stub_length.argtypes = [SEXP]
        L.stub_length.restype = C.c_ssize_t
        L.stub_string.argtypes = [SEXP, C.c_ssize_t]
        L.stub_string.restype = C.c_char_p
        L.stub_dim.argtypes = [SEXP, C.c_int]
        L.stub_errmsg.restype = C.c_char_p

    def num(self, x):
        a = np.ascontiguousarray(np.asarray(x, dtype=np.float64).reshape(-1, order="F"))
        s = self.L.stub_real(a.ctypes.data_as(C.POINTER(C.c_double)), a.size)
        if np.ndim(x) >= 2:
            d = (C.c_int * np.ndim(x))(*np.shape(x))
            self.L.stub_set_dim(s, d, np.ndim(x))
        return s

    def int(self, x):
        a = (C.c_int * len(np.atleast_1d(x)))(*[int(v) for v in np.atleast_1d(x)])
        return self.L.stub_int(a, len(a))

    def lgl(self, x):
        a = (C.c_int * len(np.atleast_1d(x)))(*[int(bool(v)) for v in np.atleast_1d(x)])
        return self.L.stub_lgl(a, len(a))

    def str(self, x):
        x = [x] if isinstance(x, str) else list(x)
        a = (C.c_char_p * max(len(x), 1))(*[v.encode() for v in x])
        return self.L.stub_str(a, len(x))

    def list(self, d):
        """list(name = value, ...): numbers become numeric(1), bools logical(1), arrays numeric matrices."""
        names = (C.c_char_p * max(len(d), 1))(*[k.encode() for k in d])
        elems = (SEXP * max(len(d), 1))(*[self.lgl(v) if isinstance(v, (bool, np.bool_)) else self.num(v) for v in d.values()])
        return self.L.stub_list(names, elems, len(d))

    def call(self, name, *args):
        """.Call(name, ...): raises with the R error message when the bridge called Rf_error."""
        fn = C.cast(getattr(self.L, name), C.c_void_p)
        a = (SEXP * len(args))(*args)
        r = self.L.stub_call(fn, len(args), a)
        assert self.L.stub_protect_depth() == 0, "PROTECT / UNPROTECT unbalanced in " + name
        msg = self.L.stub_errmsg().decode()
        if r is None:
            raise RuntimeError(msg or "R error")
        return r

    def matrix(self, s):
        n = self.L.stub_length(s)
        v = np.ctypeslib.as_array(C.cast(self.L.stub_data(s), C.POINTER(C.c_double)), shape=(max(n, 1),))[:n].copy()
        r, c = self.L.stub_dim(s, 0), self.L.stub_dim(s, 1)
        return v if r < 0 else v.reshape(r, c, order="F")

    def string(self, s):
        return self.L.stub_string(s, 0).decode()


def _configure(r, spec):
    """What buildAPT in aptb200.R does before aptR_build."""
    m = r.call("aptR_model_create", r.str(spec["code"]))
    for name, src in (spec.get("functions") or {}).items():
        r.call("aptR_model_add_function", m, r.str(name), r.str(src))
    for role, d in ((0, spec["constants"]), (1, spec["data"]), (2, spec["inits"])):
        for name, arr in d.items():
            r.call("aptR_model_set_array", m, r.str(name), r.int(role), r.num(arr))
    for s in spec["samplers"]:
        tgt = s["target"] if isinstance(s["target"], str) else "c(" + ",".join("'%s'" % t for t in s["target"]) + ")"
        r.call("aptR_model_add_sampler", m, r.str(tgt), r.str(s["type"]), r.list(s["control"]))
    r.call("aptR_model_set_monitors", m, r.int(1), r.str(spec["monitors"]), r.int(1), r.int(1))
    r.call("aptR_model_set_monitors", m, r.int(2), r.str(spec.get("monitors2") or []), r.int(1), r.int(1))
    return m


def test_bridge_marshals_model_and_errors_without_a_gpu():
    r = R()
    spec = models.logistic(300, 4, 4)  # block sampler with a propCov matrix in its control list
    m = _configure(r, spec)
    assert r.string(r.call("aptR_model_names", m, r.int(1))).split("\n") == ["beta[%d]" % (j + 1) for j in range(4)]
    # library status -> R error, after the call returned (the message is apt_last_error())
    with pytest.raises(RuntimeError, match="propCov matrix must have dimension 4x4"):  # APT:816
        bad = dict(spec["samplers"][0]["control"], propCov=np.eye(3))
        r.call("aptR_model_add_sampler", m, r.str("beta[1:4]"), r.str("sampler_RW_block_tempered"), r.list(bad))
        r.call("aptR_build", m, r.num(spec["Temps"]), r.lgl(True), r.num(1e6), r.int(1), r.int(11), r.int(0), r.lgl(False))
    with pytest.raises(RuntimeError, match="propCov must be a matrix"):
        r.call("aptR_model_add_sampler", m, r.str("beta[1:4]"), r.str("sampler_RW_block_tempered"),
               r.list(dict(temperPriors=True, propCov=np.ones(16))))
    m1 = _configure(r, spec)
    with pytest.raises(RuntimeError, match="nosuchnode"):
        r.call("aptR_model_set_monitors", m1, r.int(1), r.str(["nosuchnode"]), r.int(1), r.int(1))
        r.call("aptR_model_names", m1, r.int(1))
    m2 = _configure(r, models.multinomial_latent())  # useTempering reaches the library through the control list
    assert r.string(r.call("aptR_model_names", m2, r.int(1))).split("\n") == ["x[1]", "x[2]", "x[3]", "x[4]", "a"]
    m3 = _configure(r, models.demo_wrong_verbatim())  # the demo's nimbleFunctions travel as text (aptR_model_add_function)
    assert r.string(r.call("aptR_model_names", m3, r.int(1))).split("\n") == ["k", "theta0", "sigma0"]
    with pytest.raises(RuntimeError, match="expected 'function'"):
        r.call("aptR_model_add_function", m3, r.str("f"), r.str("42"))
    r.L.stub_reset()  # finalizers free both models


@pytest.mark.gpu
def test_bridge_builds_runs_and_returns_column_major_matrices():
    import nimbleapt_b200 as nb
    from helpers import build_engine
    r = R()
    spec = models.c1_vignette(nObs=100, K=5)
    m = _configure(r, spec)
    h = r.call("aptR_build", m, r.num(spec["Temps"]), r.lgl(True), r.num(spec["ULT"]), r.int(2), r.int(11), r.int(0), r.lgl(False))
    flags = [True, False, False, False, False, True, False]  # reset resetMV resetTempering simulateAll time adaptTemps printTemps
    nil = C.cast(r.L.stub_nil(), SEXP)
    so = r.num(np.full(2 * 100 * 2, np.nan))  # numeric(nLadders * rows * monitors): this run's rows, while it samples
    r.call("aptR_run", h, r.num(300), r.lgl(flags), r.num([10, 1]), r.int([3, -1]), so)
    ref = build_engine(spec, nLadders=2)
    ref.run(300, thin=3)
    rows = int(r.matrix(r.call("aptR_info", h, r.int(nb.api.INFO_ROWS)))[0])
    assert rows == 100
    for ladder in (0, 1):
        got = r.matrix(r.call("aptR_get", h, r.int(nb.api.APT_SAMPLES), r.int(ladder), r.num(rows), r.num(2)))
        assert got.shape == (100, 2)
        np.testing.assert_array_equal(got, ref.mvSamples.matrix(ladder))  # same seed, same ladder ids: bit-identical
        tt = r.matrix(r.call("aptR_get", h, r.int(nb.api.APT_TEMPTRAJ), r.int(ladder), r.num(rows), r.num(5)))
        np.testing.assert_array_equal(tt, ref.tempTrajLadder(ladder))
        acc = r.matrix(r.call("aptR_get", h, r.int(nb.api.APT_ACCSWAP), r.int(ladder), r.num(5), r.num(5)))
        np.testing.assert_array_equal(acc, ref.accCountSwapLadder(ladder))
        np.testing.assert_array_equal(r.matrix(so).reshape(2, 100, 2)[ladder], ref.mvSamples.matrix(ladder))
        np.testing.assert_array_equal(ref.mvSamples.matrix_column_major(ladder), ref.mvSamples.matrix(ladder))
    with pytest.raises(RuntimeError, match="output buffer too small"):  # a matrix of the wrong shape is an R error, ...
        r.call("aptR_get", h, r.int(nb.api.APT_SAMPLES), r.int(0), r.num(7), r.num(2))
    with pytest.raises(RuntimeError, match="not 300 x 2"):                 # ... not a silent short copy
        r.call("aptR_get", h, r.int(nb.api.APT_SAMPLES), r.int(0), r.num(300), r.num(2))
    # model$calculate() at two states, new inits and new data for the compiled model
    th = np.array([[3.0, 3.0], [-2.5, 0.4]])
    lp = r.matrix(r.call("aptR_logprob", h, r.num(th.T.copy())))
    tot, grp = ref.calculate(th)
    np.testing.assert_array_equal(lp[0], tot)
    np.testing.assert_array_equal(lp[1:].T, grp)
    y2 = spec["data"]["y"] + 0.25
    r.call("aptR_set_array", h, r.str("y"), r.num(y2))
    r.call("aptR_set_inits", h, r.num([1.0, -1.0]), r.int(-1))
    ref.setData("y", y2)
    ref.setInits(np.array([1.0, -1.0]))
    with pytest.raises(RuntimeError, match="one value per parameter"):
        r.call("aptR_set_inits", h, r.num([1.0, -1.0, 0.0]), r.int(-1))
    r.call("aptR_run", h, r.num(90), r.lgl(flags), r.num([10, 1]), r.int([3, -1]), nil)
    ref.run(90, thin=3)
    for ladder in (0, 1):
        got = r.matrix(r.call("aptR_get", h, r.int(nb.api.APT_SAMPLES), r.int(ladder), r.num(30), r.num(2)))
        np.testing.assert_array_equal(got, ref.mvSamples.matrix(ladder))
    # continue without reset (APT:378-380), interrupted by the user at the first poll: an R error, handle still usable
    flags[0] = False
    r.L.stub_set_interrupt(1)
    with pytest.raises(RuntimeError, match="interrupt"):
        r.call("aptR_run", h, r.num(200000), r.lgl(flags), r.num([10, 1]), r.int([3, -1]), nil)
    r.call("aptR_run", h, r.num(30), r.lgl(flags), r.num([10, 1]), r.int([3, -1]), nil)
    with pytest.raises(RuntimeError, match="niter < 0"):  # APT:350
        r.call("aptR_run", h, r.num(-5), r.lgl(flags), r.num([10, 1]), r.int([3, -1]), nil)
    with pytest.raises(RuntimeError, match="enableWAIC"):
        r.call("aptR_waic", h, r.num(0), r.int(0))
    r.L.stub_reset()  # finalizers: apt_free, apt_model_free
