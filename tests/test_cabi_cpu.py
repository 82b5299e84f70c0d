"""CPU: the C-ABI library loads, exports every symbol include/aptb200.h declares, parses/lowers/compiles models
without a GPU, and fails loudly (never silently falls back) when asked to compute without one."""
import ctypes as C
import os
import re
import numpy as np
import pytest

import models
import nimbleapt_b200 as nb
from nimbleapt_b200 import _lib
from helpers import build_engine

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_exports_match_header():
    hdr = open(os.path.join(ROOT, "include", "aptb200.h")).read()
    declared = set(re.findall(r"\b(apt_[a-z_0-9]+)\s*\(", hdr))
    declared -= {"apt_model", "apt_engine"}
    L = C.CDLL(_lib.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(L, name), "libaptb200.so does not export " + name
    assert declared == set(_lib.EXPORTS)


def test_struct_layouts_match_the_library():
    """The ctypes mirrors of the ABI structs (nimbleapt_b200/_lib.py) have the sizes the library was compiled with."""
    import ctypes as C
    from nimbleapt_b200 import _lib
    L = _lib.lib()
    for which, cls in ((0, _lib.SamplerControl), (1, _lib.BuildOpts), (2, _lib.RunArgs), (3, _lib.Timing)):
        assert L.apt_abi_sizeof(which) == C.sizeof(cls), cls.__name__
    assert L.apt_abi_sizeof(99) == -1
    ra = _lib.RunArgs()
    L.apt_run_args_default(C.byref(ra))
    assert (ra.reset, ra.adaptTemps, ra.thin, ra.thin2, ra.tuneTemper1, ra.tuneTemper2) == (1, 1, -1, -1, 10.0, 1.0)  # APT:336-348
    assert not ra.samples_out and ra.samples_out_capacity == 0


def test_lowering_emits_expected_cuda():
    spec = models.c4_glmm(G=20, per=10, K=4)
    m = nb.nimbleModel(spec["code"], spec["constants"], spec["data"], spec["inits"])
    h = C.c_void_p()
    L = _lib.lib()
    _lib.check(L.apt_model_create(spec["code"].encode(), C.byref(h)))
    L.apt_model_free(h)
    with pytest.raises(nb.AptError, match="parse error"):
        nb.nimbleModel("{ y ~ dnorm(0, 1 }")
    with pytest.raises(nb.AptError) as ei:
        build_engine(dict(spec, code=spec["code"].replace("dpois", "dweib")))
    assert "unsupported distribution" in str(ei.value)


def test_no_gpu_means_loud_failure():
    import torch  # only to learn whether this host has a GPU
    if torch.cuda.is_available():
        pytest.skip("host has a GPU")
    with pytest.raises(nb.AptError, match="no CUDA device|no CPU fallback"):
        build_engine(models.c2_mixture(K=4))


def test_control_validation_mirrors_reference():
    spec = models.c2_mixture(K=4)
    m = nb.nimbleModel(spec["code"], spec["constants"], spec["data"], spec["inits"])
    conf = nb.configureMCMC(m, monitors="theta")
    conf.addSampler("theta[1]", type="sampler_RW_tempered", control={})
    with pytest.raises(nb.AptError, match="temperPriors unspecified in sampler_RW_tempered"):  # APT:660
        nb.buildAPT(conf, Temps=[1, 2, 3])
    conf.removeSamplers()
    conf.addSampler("theta[1:2]", type="sampler_RW_tempered", control=dict(temperPriors=True))
    with pytest.raises(nb.AptError, match="more than one target"):  # APT:682
        nb.buildAPT(conf, Temps=[1, 2, 3])
    conf.removeSamplers()
    conf.addSampler("theta[1]", type="sampler_RW_tempered", control=dict(temperPriors=True, log=True, reflective=True))
    with pytest.raises(nb.AptError, match="reflective RW sampler on a log scale"):  # APT:684
        nb.buildAPT(conf, Temps=[1, 2, 3])
    conf.removeSamplers()
    conf.addSampler("theta[1:2]", type="sampler_RW_block_tempered", control=dict(temperPriors=True, propCov=np.eye(3)))
    with pytest.raises(nb.AptError, match="propCov matrix must have dimension 2x2"):  # APT:816
        nb.buildAPT(conf, Temps=[1, 2, 3])
    conf.removeSamplers()
    conf.addSampler("theta[1:2]", type="sampler_RW_block_tempered", control=dict(temperPriors=True, propCov=np.array([[1, 0.5], [0.1, 1]])))
    with pytest.raises(nb.AptError, match="propCov matrix must be symmetric"):  # APT:817
        nb.buildAPT(conf, Temps=[1, 2, 3])
    with pytest.raises(ValueError, match="unsupported sampler type"):
        conf.addSampler("theta[1]", type="sampler_AF_slice", control=dict(temperPriors=True))
    conf.removeSamplers()
    conf.addSampler("theta[1:2]", type="sampler_slice_tempered", control={})
    with pytest.raises(nb.AptError, match="cannot use slice sampler on more than one target node"):  # APT:928
        nb.buildAPT(conf, Temps=[1, 2, 3])
    with pytest.raises(TypeError, match="conf must either be a nimbleModel or a MCMCconf"):  # APT:252
        nb.buildAPT("not a conf", Temps=[1, 2])


def test_multinomial_sampler_lowers_and_setup_checks():
    """sampler_RW_multinomial_tempered (APT:1005-1145): lowering of a dmulti target and the setup code's checks."""
    spec = models.multinomial_latent()
    apt = build_engine(spec, compileOnly=True)
    src = apt.lowered_source()
    assert "apt::multinomial_update<Model, S0>" in src and "apt::dmulti_log<4>" in src and "NTOTAL = 30" in src
    assert "blk_keep(int i) { return i == 0 ||" in src  # sampler$reset() leaves u alone (APT:1133-1142)
    assert os.path.exists(apt.modulePath)

    def bad(samplers, **kw):
        b = dict(spec)
        b["samplers"] = samplers
        return build_engine(b, compileOnly=True, **kw)
    with pytest.raises(ValueError, match="unknown control option"):
        bad([dict(target="x[1:4]", type="sampler_RW_multinomial_tempered", control=dict(temperPriors=True))])
    with pytest.raises(nb.AptError, match="useTempering unspecified in sampler_RW_multinomial_tempered"):  # APT:1015
        bad([dict(target="x[1:4]", type="sampler_RW_multinomial_tempered", control={})])
    with pytest.raises(nb.AptError, match="can only use RW_multinomial sampler for multinomial distributions"):  # APT:1046
        bad([dict(target="a", type="sampler_RW_multinomial_tempered", control=dict(useTempering=True))])
    with pytest.raises(nb.AptError, match="adaptInterval < 100 is not recommended"):  # APT:1048
        bad([dict(target="x[1:4]", type="sampler_RW_multinomial_tempered", control=dict(useTempering=True, adaptInterval=50))])
    with pytest.raises(nb.AptError, match="whole multinomial node"):
        bad([dict(target="x[1:3]", type="sampler_RW_multinomial_tempered", control=dict(useTempering=True))])
    with pytest.raises(nb.AptError, match="grid engine: sampler_RW_multinomial_tempered is not supported"):
        bad(spec["samplers"], engine=2)


def test_chains_per_tile_lowering(monkeypatch):
    """G::MULTI (lower.cpp::choose_shape): many rungs over a shared Bernoulli-logit likelihood -> two chains per tile,
    generated eval_multi with the short inner product unrolled completely; off by switch, off for few rungs, off for the
    grid engine."""
    spec = models.c5_small_logistic(N=602, p=6, K=16)
    src = build_engine(spec, nLadders=4, compileOnly=True).lowered_source()
    assert "MULTI = 2" in src and "apt::rw_update_multi<Model, S0, MULTI>(c, D, k)" in src
    assert "eval_multi(const Data& D, const double* const (&thv)[CH]" in src and "apt::bern_logit_tile<4, CH>(y_, e_, a_)" in src
    assert '_Pragma("unroll")\n for (int e' in src and '_Pragma("unroll 5")' not in src
    monkeypatch.setenv("APTB200_MULTI", "0")
    src0 = build_engine(spec, nLadders=4, compileOnly=True).lowered_source()
    assert "MULTI = 1" in src0 and "rw_update_multi" not in src0 and '_Pragma("unroll 5")' in src0
    monkeypatch.delenv("APTB200_MULTI")
    few = build_engine(models.c5_small_logistic(N=602, p=6, K=4), nLadders=64, compileOnly=True).lowered_source()
    assert "MULTI = 1" in few  # 4 rungs x 32 lanes fit one CTA: nothing to share
    grid = build_engine(spec, nLadders=1, compileOnly=True, engine=2).lowered_source()
    assert "MULTI = 1" in grid and "ENGINE = 2" in grid


def test_demo_model_lowers_with_slice_sampler():
    """demo/APT_demo.R:83-92,173-177: dcat, a constant array indexed by the sampled node, sampler_slice_tempered."""
    apt = build_engine(models.demo_wrong(), compileOnly=True)
    src = apt.lowered_source()
    assert "apt::slice_update<Model, S2>" in src and "DISCRETE = true" in src
    assert "apt::dcat_log<10>" in src and "apt::dyn_index(th[0], 10)" in src
    assert os.path.exists(apt.modulePath)
    with pytest.raises(nb.AptError, match="grid engine: sampler_slice_tempered is not supported"):
        build_engine(models.demo_wrong(), compileOnly=True, engine=2)


def test_concurrent_cold_builds_share_one_cache_entry(tmp_path):
    """Several processes (8 ranks calling buildAPT at once) compile the same model into an empty cache whose path contains a
    space: every one of them succeeds (sources, logs and modules are written under process-unique names and renamed into
    place, shell arguments are quoted) and one module remains."""
    import subprocess
    import sys
    cache = tmp_path / "jit cache"
    cache.mkdir()
    code = ("import sys; sys.path[:0] = [%r, %r]\n"
            "import models\nfrom helpers import build_engine\n"
            "apt = build_engine(models.c2_mixture(n=6, K=4, tmax=20.0), compileOnly=True)\nprint(apt.modulePath)\n") % (
                os.path.dirname(os.path.dirname(os.path.abspath(__file__))), os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, APTB200_CACHE=str(cache))
    procs = [subprocess.Popen([sys.executable, "-c", code], env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True) for _ in range(3)]
    outs = [p.communicate(timeout=600) for p in procs]
    assert all(p.returncode == 0 for p in procs), [o[1][-400:] for o in outs]
    paths = {o[0].strip().splitlines()[-1] for o in outs}
    assert len(paths) == 1 and os.path.exists(paths.pop())
    left = sorted(os.listdir(cache))
    assert [f for f in left if f.endswith(".so")] and not [f for f in left if ".tmp." in f], left
