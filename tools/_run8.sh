mkdir -p gpurun_out
tools/gpu_variants.sh dev ev1 ev2 dev ev1 2>&1 | tee gpurun_out/t09_variants.log
