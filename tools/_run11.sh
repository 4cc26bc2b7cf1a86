mkdir -p gpurun_out
{ echo default; python tools/run_c4.py 2>&1 | tail -2; echo vw6; LYON_CUDA_LIB=$PWD/lyon_b200/liblyon_cuda_vw6.so python tools/run_c4.py 2>&1 | tail -2; } | tee gpurun_out/y03_c4.log
