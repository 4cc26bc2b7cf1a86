mkdir -p gpurun_out
export LYON_CUDA_LIB=$PWD/lyon_b200/liblyon_cuda_rev.so
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/w02_tests.log 2>&1
tail -2 gpurun_out/w02_tests.log
tools/gpu_variants.sh ch7 ch8 rev ch7 ch8 rev | tee gpurun_out/w02_variants.log
