mkdir -p gpurun_out
export LYON_CUDA_LIB=$PWD/lyon_b200/liblyon_cuda_dev.so
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29541 tools/sharded_check.py 2>&1 | tail -5 | tee gpurun_out/t10_sharded.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu 2>&1 | tail -2 | tee -a gpurun_out/t10_sharded.log
