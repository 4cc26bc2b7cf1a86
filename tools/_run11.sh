mkdir -p gpurun_out
export LYON_CUDA_LIB=$PWD/lyon_b200/liblyon_cuda_tpf.so
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/x05_tests.log 2>&1
tail -2 gpurun_out/x05_tests.log
tools/gpu_variants.sh nospf spf tpf nospf spf tpf | tee gpurun_out/x05_variants.log
