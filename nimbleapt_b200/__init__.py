"""nimbleapt_b200 — B200-native adaptive parallel tempering behind nimbleAPT's interface.

Product code only: the C ABI library (libaptb200.so: BUGS-subset lowering + nvcc driver + device engine modules)
and the host-side mirror of the reference's R interface.  Nothing here imports `oracle/` (test infrastructure).
"""
from .api import (APT, AptError, MCMCconf, NimbleModel, as_matrix, buildAPT, compileNimble, configureMCMC, nimbleCode,
                  nimbleFunction, nimbleModel, plotTempTraj, runAPT, sampler_RW_block_tempered, sampler_RW_multinomial_tempered,
                  sampler_RW_tempered, sampler_slice_tempered)

__all__ = ["APT", "AptError", "MCMCconf", "NimbleModel", "as_matrix", "buildAPT", "compileNimble", "configureMCMC",
           "nimbleCode", "nimbleFunction", "nimbleModel", "plotTempTraj", "runAPT", "sampler_RW_block_tempered",
           "sampler_RW_multinomial_tempered", "sampler_RW_tempered", "sampler_slice_tempered"]
