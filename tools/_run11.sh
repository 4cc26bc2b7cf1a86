mkdir -p gpurun_out
export LYON_CUDA_LIB=$PWD/lyon_b200/liblyon_cuda_chp.so
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/x03_tests.log 2>&1
tail -2 gpurun_out/x03_tests.log
tools/gpu_variants.sh er0 chp er0 chp | tee gpurun_out/x03_variants.log
