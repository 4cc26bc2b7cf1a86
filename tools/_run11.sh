mkdir -p gpurun_out
export LYON_CUDA_LIB=$PWD/lyon_b200/liblyon_cuda_peel.so
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/w03_tests.log 2>&1
tail -2 gpurun_out/w03_tests.log
tools/gpu_variants.sh nopeel peel nopeel peel | tee gpurun_out/w03_variants.log
