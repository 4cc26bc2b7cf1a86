mkdir -p gpurun_out
export LYON_CUDA_LIB=$PWD/lyon_b200/liblyon_cuda_dev.so
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/u01_tests.log 2>&1
tail -3 gpurun_out/u01_tests.log
tools/gpu_variants.sh base dev base dev | tee gpurun_out/u01_variants.log
LYON_PIPE=ctile LYON_CUDA_LIB=$PWD/lyon_b200/liblyon_cuda_phase.so timeout 600 python tools/phase_profile.py 200000 2>&1 | tee gpurun_out/u01_phase.log
timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_stroke_tile -s 1 -c 1 -o gpurun_out/u01_tile -f python tools/run_c3.py 200000 2 > gpurun_out/u01_ncu.log 2>&1
ncu -i gpurun_out/u01_tile.ncu-rep --page raw --csv > gpurun_out/u01_raw.csv 2>/dev/null
ncu -i gpurun_out/u01_tile.ncu-rep --page source --csv > gpurun_out/u01_src.csv 2>/dev/null
tail -2 gpurun_out/u01_ncu.log
