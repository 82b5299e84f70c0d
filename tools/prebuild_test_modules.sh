#!/bin/bash
# Fill nimbleapt_b200/_jit_cache with the lowered-model modules the GPU tests build, on a machine without a GPU
# (nvcc cross-compiles for sm_100a), so that the GPU box does not spend its time in nvcc.  Every test "fails" at
# its first buildAPT by design (APTB200_PREBUILD); tests that build a second, different model leave that one to
# the GPU box.
cd "$(dirname "$0")/.."
APTB200_PREBUILD=1 python -m pytest tests -m gpu -q -n ${JOBS:-8} -p no:cacheprovider 2>&1 | tail -3
ls nimbleapt_b200/_jit_cache/*.so | wc -l
