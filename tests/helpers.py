"""Shared builders: the same workload spec (tests/models.py) drives the CUDA engine (through the C ABI) and the
CPU oracle (test infrastructure)."""
import numpy as np

import nimbleapt_b200 as nb
from oracle.oracle import OracleAPT


def build_engine(spec, nLadders=1, **kw):
    m = nb.nimbleModel(nb.nimbleCode(spec["code"]), constants=spec["constants"], data=spec["data"], inits=spec["inits"])
    conf = nb.configureMCMC(m, monitors=spec["monitors"], monitors2=spec.get("monitors2"))
    conf.removeSamplers()
    for s in spec["samplers"]:
        conf.addSampler(s["target"], type=s["type"], control=s["control"])
    return nb.buildAPT(conf, Temps=spec["Temps"], ULT=spec["ULT"], nLadders=nLadders, **kw)


def build_oracle(spec, nLadders=1, seed=11, ladder0=0):
    return OracleAPT(spec["code"], spec["constants"], spec["data"], spec["inits"], spec["samplers"], spec["monitors"],
                     spec.get("monitors2", ()), spec["Temps"], spec["ULT"], True, nLadders, seed, ladder0)


def make_tables(niter, L, K, NU, DMAX, seed=1):
    rng = np.random.default_rng(seed)
    Z = rng.standard_normal((niter, L, K, NU, DMAX))
    U = rng.random((niter, L, K, NU))
    S = rng.random((niter, L, K, 2))
    return Z, U, S


def run_pair(spec, niter, L=1, inject=True, seed=11, thin=1, record=True, build_kw=None, **run_kw):
    """Run engine and oracle on identical inputs; returns (engine, oracle, records)."""
    eng = build_engine(spec, nLadders=L, seed=seed, record=record, **(build_kw or {}))
    orc = build_oracle(spec, nLadders=L, seed=seed)
    K, NU, DMAX = eng.nTemps, eng.info(nb.api.INFO_N_UNITS), eng.info(nb.api.INFO_DMAX)
    recs = []
    inj = None
    if inject:
        Z, U, S = make_tables(niter, L, K, NU, DMAX, seed=seed + 100)
        inj = dict(Z=Z, U=U, S=S)
    for l in range(L):
        if inject:
            recs.append(orc.set_tables(l, Z[:, l].copy(), U[:, l].copy(), S[:, l].copy(), record=True, niter=niter))
        else:
            recs.append(orc.set_tables(l, record=True, niter=niter))
    eng.run(niter, thin=thin, _inject=inj, **run_kw)
    orc.run(niter, thin=thin, **run_kw)
    return eng, orc, recs
