mkdir -p gpurun_out
export LYON_CUDA_LIB=$PWD/lyon_b200/liblyon_cuda_pp.so
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/v05_tests.log 2>&1
tail -2 gpurun_out/v05_tests.log
{ tools/gpu_variants.sh pp d8 pp d8
LYON_PREPASS=warp tools/gpu_variants.sh pp
tools/gpu_variants.sh nt320:320,3520,0 nt288:288,3520,0 nt320:320,3520,0 nt288:288,3520,0
} | tee gpurun_out/v05_variants.log
