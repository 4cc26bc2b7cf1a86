mkdir -p gpurun_out
export LYON_CUDA_LIB=$PWD/lyon_b200/liblyon_cuda_f2.so
timeout 900 python -m pytest tests -m gpu -x -q -k "flatten or c2 or measure or abi" > gpurun_out/v06_tests.log 2>&1
tail -3 gpurun_out/v06_tests.log
{ for t in f2 f2w512 f2w512b6; do echo -n "$t: "; LYON_CUDA_LIB=$PWD/lyon_b200/liblyon_cuda_$t.so python tools/run_c2.py 2>&1 | tail -1; done
echo -n "f2 block: "; LYON_FLAT=block LYON_CUDA_LIB=$PWD/lyon_b200/liblyon_cuda_f2.so python tools/run_c2.py 2>&1 | tail -1
} | tee gpurun_out/v06_flat.log
