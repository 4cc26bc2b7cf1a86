mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "verb_round or host_pipeline or sharded" > gpurun_out/y01_tests.log 2>&1
tail -3 gpurun_out/y01_tests.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29541 tools/sharded_check.py 2>&1 | tail -3 | tee gpurun_out/y01_sharded.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu 2>&1 | tail -1 > gpurun_out/y01_bench_n2.json
python -c "
import json
d=json.loads(open('gpurun_out/y01_bench_n2.json').read())
print('N=2 ms', d['ms_per_step'], 'Gvert/s', d['value']/1e9, 'e2e', d['e2e']['ms_per_step'] if d.get('e2e') else None)
"
