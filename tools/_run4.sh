mkdir -p gpurun_out
export LYON_CUDA_LIB=$PWD/lyon_b200/liblyon_cuda_dev.so
timeout 600 python tools/wt_smoke.py 200 v 2>&1 | tail -30 > gpurun_out/s04_smoke.log
timeout 600 env LYON_PIPE=wtile python tools/wt_smoke.py 200 v 2>&1 | tail -30 >> gpurun_out/s04_smoke.log
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/s04_tests.log 2>&1
{
for pipe in ctile wtile; do
  LYON_PIPE=$pipe timeout 600 python bench.py --no-e2e --no-cpu --no-flatten --steps 5 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$pipe', 'ms %.3f frac %.4f kernel_ms %s' % (d['ms_per_step'], d['roofline']['frac'], d['roofline'].get('kernel_ms')))"
done
LYON_PIPE=ctile LYON_CUDA_LIB=$PWD/lyon_b200/liblyon_cuda_phase.so timeout 600 python tools/phase_profile.py 200000
python tools/run_c4.py 2>&1 | tail -4
} > gpurun_out/s04_bench.log 2>&1
