mkdir -p gpurun_out
export LYON_CUDA_LIB=$PWD/lyon_b200/liblyon_cuda_dev.so
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/t13_tests.log 2>&1
tail -4 gpurun_out/t13_tests.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 3 python tools/sanitize_run.py 150 > gpurun_out/t13_racecheck.log 2>&1; echo "racecheck rc $?"; tail -3 gpurun_out/t13_racecheck.log
timeout 900 python bench.py --no-cpu --no-e2e --steps 5 2>&1 | tail -1 > gpurun_out/t13_bench.json
python - <<'PY'
import json
d=json.load(open('gpurun_out/t13_bench.json'))
print('C3 ms', d['ms_per_step'], 'frac', d['roofline']['frac'])
print('flatten', d['flatten']['ms_per_step'], d['flatten']['value']/1e9, 'Gseg/s frac', d['flatten']['roofline']['frac'], 'kernel_ms', d['flatten']['roofline']['kernel_ms'])
for k,v in d['configs'].items():
    print(k, {kk: vv for kk, vv in v.items() if kk in ('ms','value','sweep','paths_per_s')} if k!='C5_tolerance_sweep' else [(s['tolerance'], round(s['ms'],3), s['vertices']) for s in v['sweep']])
PY
