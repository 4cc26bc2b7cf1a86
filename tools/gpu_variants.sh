#!/bin/bash
# bench (device-resident stroke only) for several dev builds: tools/gpu_variants.sh tag[:LYON_TILE] ...
for spec in "$@"; do
  t=${spec%%:*}; tile=""; [[ "$spec" == *:* ]] && tile=${spec#*:}
  LYON_TILE=$tile LYON_CUDA_LIB=$PWD/lyon_b200/liblyon_cuda_$t.so python bench.py --no-e2e --no-cpu --no-flatten --steps 5 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$spec', 'Gvert/s %.3f ms %.3f frac %.4f' % (d['value']/1e9, d['ms_per_step'], d['roofline']['frac']))"
done
