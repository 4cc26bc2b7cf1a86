mkdir -p gpurun_out
export LYON_CUDA_LIB=$PWD/lyon_b200/liblyon_cuda_dev.so
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/t11_tests.log 2>&1
tail -15 gpurun_out/t11_tests.log
timeout 900 python bench.py --no-cpu --no-flatten --steps 3 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('ms', d['ms_per_step'], 'e2e', d['e2e'])" | tee gpurun_out/t11_bench.log
