"""ctypes binding of libaptb200.so (include/aptb200.h).  Fails loudly: no fallback of any kind."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libaptb200.so")


class SamplerControl(C.Structure):
    _fields_ = [("log", C.c_int), ("reflective", C.c_int), ("adaptive", C.c_int), ("adaptScaleOnly", C.c_int),
                ("adaptInterval", C.c_int), ("adaptFactorExponent", C.c_double), ("scale", C.c_double),
                ("temperPriors", C.c_int), ("propCov", C.POINTER(C.c_double)), ("propCov_dim", C.c_int),
                ("width", C.c_double), ("sliceMaxSteps", C.c_int), ("useTempering", C.c_int)]


class BuildOpts(C.Structure):
    _fields_ = [("monitorTmax", C.c_int), ("ULT", C.c_double), ("n_ladders", C.c_int), ("ladder0", C.c_uint32),
                ("seed", C.c_uint32), ("device", C.c_int), ("record_decisions", C.c_int), ("fuse_links", C.c_int),
                ("lanes_per_chain", C.c_int), ("cta_threads", C.c_int), ("engine", C.c_int),
                ("rungs_per_pass", C.c_int), ("verbose", C.c_int), ("cluster_ctas", C.c_int), ("enable_waic", C.c_int)]


ABORT_FN = C.CFUNCTYPE(C.c_int, C.c_void_p)


class RunArgs(C.Structure):
    _fields_ = [("niter", C.c_int64), ("reset", C.c_int), ("resetMV", C.c_int), ("resetTempering", C.c_int),
                ("simulateAll", C.c_int), ("time", C.c_int), ("adaptTemps", C.c_int), ("printTemps", C.c_int),
                ("tuneTemper1", C.c_double), ("tuneTemper2", C.c_double), ("progressBar", C.c_int),
                ("thin", C.c_int), ("thin2", C.c_int), ("should_abort", ABORT_FN), ("abort_arg", C.c_void_p),
                ("iters_per_launch", C.c_int), ("inject_normals", C.POINTER(C.c_double)),
                ("inject_accept_u", C.POINTER(C.c_double)), ("inject_swap_u", C.POINTER(C.c_double)),
                ("samples_out", C.POINTER(C.c_double)), ("samples_out_capacity", C.c_int64),
                ("gather", C.c_int), ("gather_root", C.c_int)]


class Timing(C.Structure):
    _fields_ = [("kernel_ms", C.c_double), ("main_kernel_ms", C.c_double), ("launches", C.c_int64),
                ("main_launches", C.c_int64), ("run_wall_ms", C.c_double), ("gather_tail_ms", C.c_double)]


# every symbol include/aptb200.h declares
EXPORTS = ["apt_last_error", "apt_version", "apt_abi_sizeof", "apt_sampler_control_default", "apt_build_opts_default",
           "apt_run_args_default", "apt_model_create", "apt_model_free", "apt_model_set_array",
           "apt_model_add_sampler", "apt_model_add_function", "apt_model_remove_samplers", "apt_model_set_monitors", "apt_model_set_thin",
           "apt_model_names", "apt_model_lower", "apt_model_compile", "apt_build", "apt_free", "apt_run", "apt_info", "apt_get", "apt_get_layout",
           "apt_get_timing", "apt_logprob", "apt_waic", "apt_set_inits", "apt_set_array", "apt_comm_unique_id", "apt_comm_init",
           "apt_comm_allgather_samples", "apt_comm_gather_samples", "apt_comm_allreduce_sum", "apt_comm_finalize"]

_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError("libaptb200.so is missing at %s: run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "or `make -C nimbleapt_b200/csrc` (there is no fallback implementation)" % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    dp, vp = C.POINTER(C.c_double), C.c_void_p
    L.apt_last_error.restype = C.c_char_p
    L.apt_version.restype = C.c_char_p
    L.apt_abi_sizeof.restype = C.c_int64
    L.apt_abi_sizeof.argtypes = [C.c_int]
    # the ctypes mirrors above must have the layout the library was compiled with
    for which, cls in ((0, SamplerControl), (1, BuildOpts), (2, RunArgs), (3, Timing)):
        if L.apt_abi_sizeof(which) != C.sizeof(cls):
            raise AptError("ABI mismatch: %s is %d bytes here, %d in libaptb200.so (rebuild nimbleapt_b200/csrc)"
                           % (cls.__name__, C.sizeof(cls), L.apt_abi_sizeof(which)))
    L.apt_sampler_control_default.argtypes = [C.POINTER(SamplerControl)]
    L.apt_build_opts_default.argtypes = [C.POINTER(BuildOpts)]
    L.apt_run_args_default.argtypes = [C.POINTER(RunArgs)]
    L.apt_model_create.argtypes = [C.c_char_p, C.POINTER(vp)]
    L.apt_model_free.argtypes = [vp]
    L.apt_model_set_array.argtypes = [vp, C.c_char_p, C.c_int, dp, C.POINTER(C.c_int64), C.c_int]
    L.apt_model_add_sampler.argtypes = [vp, C.c_char_p, C.c_char_p, C.POINTER(SamplerControl)]
    L.apt_model_add_function.argtypes = [vp, C.c_char_p, C.c_char_p]
    L.apt_model_remove_samplers.argtypes = [vp]
    L.apt_model_set_monitors.argtypes = [vp, C.c_int, C.POINTER(C.c_char_p), C.c_int]
    L.apt_model_set_thin.argtypes = [vp, C.c_int, C.c_int]
    L.apt_model_names.argtypes = [vp, C.c_int, C.c_char_p, C.c_int64]
    L.apt_model_lower.argtypes = [vp, C.c_int, C.POINTER(BuildOpts), C.c_char_p, C.c_int64, C.POINTER(C.c_int64)]
    L.apt_model_compile.argtypes = [vp, C.c_int, C.POINTER(BuildOpts), C.c_char_p, C.c_int64]
    L.apt_build.argtypes = [vp, dp, C.c_int, C.POINTER(BuildOpts), C.POINTER(vp)]
    L.apt_free.argtypes = [vp]
    L.apt_run.argtypes = [vp, C.POINTER(RunArgs)]
    L.apt_info.restype = C.c_int64
    L.apt_info.argtypes = [vp, C.c_int]
    L.apt_get.argtypes = [vp, C.c_int, C.c_int, dp, C.c_int64, C.POINTER(C.c_int64)]
    L.apt_get_layout.argtypes = [vp, C.c_int, C.c_int, C.c_int, dp, C.c_int64, C.POINTER(C.c_int64)]
    L.apt_get_timing.argtypes = [vp, C.POINTER(Timing)]
    L.apt_logprob.argtypes = [vp, dp, C.c_int, dp, dp]
    L.apt_waic.argtypes = [vp, C.c_int64, C.c_int, dp]
    L.apt_set_inits.argtypes = [vp, dp, C.c_int]
    L.apt_set_array.argtypes = [vp, C.c_char_p, dp, C.c_int64]
    L.apt_comm_unique_id.argtypes = [C.POINTER(C.c_ubyte)]
    L.apt_comm_init.argtypes = [vp, C.c_int, C.c_int, C.POINTER(C.c_ubyte)]
    L.apt_comm_allgather_samples.argtypes = [vp, C.c_int, dp, C.c_int64, C.POINTER(C.c_int64)]
    L.apt_comm_gather_samples.argtypes = [vp, C.c_int, C.c_int, dp, C.c_int64, C.POINTER(C.c_int64)]
    L.apt_comm_allreduce_sum.argtypes = [vp, dp, C.c_int64]
    L.apt_comm_finalize.argtypes = [vp]
    _lib = L
    return L


class AptError(RuntimeError):
    pass


def check(rc):
    if rc != 0:
        raise AptError(lib().apt_last_error().decode("utf-8", "replace"))
