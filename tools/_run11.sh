mkdir -p gpurun_out
export LYON_CUDA_LIB=$PWD/lyon_b200/liblyon_cuda_pack.so
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/w05_tests.log 2>&1
tail -2 gpurun_out/w05_tests.log
tools/gpu_variants.sh packb pack packb pack | tee gpurun_out/w05_variants.log
for t in packb pack; do LYON_CUDA_LIB=$PWD/lyon_b200/liblyon_cuda_$t.so ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_tile_pack -c 6 --csv python tools/run_c3.py 1000000 2 2>/dev/null | grep k_tile_pack | awk -F'","' -v t=$t '{print t, $NF}' ; done | tee gpurun_out/w05_pack_ns.log
