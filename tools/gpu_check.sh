#!/bin/bash
# On the GPU box: parity tests + short bench with the dev build(s), then the phase profile with the instrumented build.
export LYON_CUDA_LIB=$PWD/lyon_b200/liblyon_cuda_dev.so
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
tools/gpu_variants.sh dev "$@"
LYON_CUDA_LIB=$PWD/lyon_b200/liblyon_cuda_phase.so python tools/phase_profile.py 200000 2>&1 | tail -14
