mkdir -p gpurun_out
for t in f128w3584 f128w2560 f64w1792 f192w4608 f128w3072; do echo -n "$t: "; LYON_CUDA_LIB=$PWD/lyon_b200/liblyon_cuda_$t.so python tools/run_c2.py 2>&1 | tail -1; done | tee gpurun_out/t16_flat_variants.log
tools/gpu_variants.sh base nopf dev base nopf dev | tee gpurun_out/t16_bench.log
