mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/x01_tests.log 2>&1
tail -3 gpurun_out/x01_tests.log
timeout 900 python bench.py > gpurun_out/x01_bench.json 2> gpurun_out/x01_bench.err
tail -c 300 gpurun_out/x01_bench.json
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/x01_bench_ref.json 2> gpurun_out/x01_bench_ref.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/x01_launches.csv python bench.py --no-cpu --no-e2e --steps 2 --warmup 1 > gpurun_out/x01_ncu_bench.log 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_stroke_tile -s 1 -c 1 -o gpurun_out/x01_tile -f python tools/run_c3.py 200000 2 > gpurun_out/x01_ncu.log 2>&1
ncu -i gpurun_out/x01_tile.ncu-rep --page raw --csv > gpurun_out/x01_raw.csv 2>/dev/null
ncu -i gpurun_out/x01_tile.ncu-rep --page source --csv > gpurun_out/x01_src.csv 2>/dev/null
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/x01_smoke.log 2>&1; tail -2 gpurun_out/x01_smoke.log
