mkdir -p gpurun_out
export LYON_CUDA_LIB=$PWD/lyon_b200/liblyon_cuda_pk.so
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/u04_tests.log 2>&1
tail -3 gpurun_out/u04_tests.log
tools/gpu_variants.sh dev pk pk_plain pk_i64 dev pk pk_plain pk_i64 | tee gpurun_out/u04_variants.log
