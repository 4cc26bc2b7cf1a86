mkdir -p gpurun_out
export LYON_CUDA_LIB=$PWD/lyon_b200/liblyon_cuda_dev.so
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/t08_tests.log 2>&1
tail -5 gpurun_out/t08_tests.log
unset LYON_CUDA_LIB
tools/gpu_variants.sh dev "$@" 2>&1 | tee gpurun_out/t08_bench.log
LYON_CUDA_LIB=$PWD/lyon_b200/liblyon_cuda_phase.so LYON_PIPE=ctile timeout 600 python tools/phase_profile.py 200000 2>&1 | tee -a gpurun_out/t08_bench.log
LYON_CUDA_LIB=$PWD/lyon_b200/liblyon_cuda_dev.so timeout 900 python bench.py --no-cpu --no-flatten --steps 3 2>&1 | tail -1 | tee -a gpurun_out/t08_bench.log
