mkdir -p gpurun_out
for t in pm1 pm4 pm5 pm6; do LYON_CUDA_LIB=$PWD/lyon_b200/liblyon_cuda_$t.so ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_path_slots -c 3 --csv python tools/run_c3.py 1000000 2 2>/dev/null | grep k_path_slots | awk -F'","' -v t=$t '{print t, $NF}' ; done | tee gpurun_out/y04_prepass_ns.log
tools/gpu_variants.sh pm1 pm4 pm5 pm1 pm4 pm5 | tee gpurun_out/y04_variants.log
