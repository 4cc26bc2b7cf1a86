mkdir -p gpurun_out
export LYON_CUDA_LIB=$PWD/lyon_b200/liblyon_cuda_mg.so
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/v02_tests.log 2>&1
tail -2 gpurun_out/v02_tests.log
tools/gpu_variants.sh mg0 mg mg0 mg | tee gpurun_out/v02_variants.log
