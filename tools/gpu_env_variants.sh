#!/bin/bash
# bench (device-resident stroke only) of one dev build under several environment settings:
#   tools/gpu_env_variants.sh tag "VAR=val VAR2=val" "VAR=val" ...
t=$1; shift
for envs in "$@"; do
  env $envs LYON_CUDA_LIB=$PWD/lyon_b200/liblyon_cuda_$t.so python bench.py --no-e2e --no-cpu --no-flatten --steps 5 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$envs', 'Gvert/s %.3f ms %.3f frac %.4f' % (d['value']/1e9, d['ms_per_step'], d['roofline']['frac']))"
done
