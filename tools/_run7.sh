mkdir -p gpurun_out
export LYON_CUDA_LIB=$PWD/lyon_b200/liblyon_cuda_dev.so
timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/t06_tests.log 2>&1
tail -3 gpurun_out/t06_tests.log
unset LYON_CUDA_LIB
tools/gpu_variants.sh dev "$@" 2>&1 | tee gpurun_out/t06_bench.log
LYON_CUDA_LIB=$PWD/lyon_b200/liblyon_cuda_phase.so LYON_PIPE=ctile timeout 600 python tools/phase_profile.py 200000 2>&1 | tee -a gpurun_out/t06_bench.log
