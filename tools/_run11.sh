mkdir -p gpurun_out
export LYON_CUDA_LIB=$PWD/lyon_b200/liblyon_cuda_ppf.so
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/x06_tests.log 2>&1
tail -2 gpurun_out/x06_tests.log
tools/gpu_variants.sh nopre pre noppf ppf nopre pre noppf ppf | tee gpurun_out/x06_variants.log
