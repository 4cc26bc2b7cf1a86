mkdir -p gpurun_out
for t in em1 em6 em8; do LYON_CUDA_LIB=$PWD/lyon_b200/liblyon_cuda_$t.so ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_stroke_ends -c 3 --csv python tools/run_c3.py 1000000 2 2>/dev/null | grep k_stroke_ends | awk -F'","' -v t=$t '{print t, $NF}' ; done | tee gpurun_out/y05_ends_ns.log
