mkdir -p gpurun_out
tools/gpu_variants.sh vu0 vu1 vu2 vu8 vu0 vu1 vu2 vu8 | tee gpurun_out/y02_variants.log
