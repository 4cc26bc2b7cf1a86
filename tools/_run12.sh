mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/w01_tests.log 2>&1
tail -3 gpurun_out/w01_tests.log
timeout 900 python bench.py > gpurun_out/w01_bench.json 2> gpurun_out/w01_bench.err
tail -c 300 gpurun_out/w01_bench.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/w01_bench_ref.json 2> gpurun_out/w01_bench_ref.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/w01_launches.csv python bench.py --no-cpu --no-e2e --steps 2 --warmup 1 > gpurun_out/w01_ncu_bench.log 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_stroke_tile -s 1 -c 1 -o gpurun_out/w01_tile -f python tools/run_c3.py 200000 2 > gpurun_out/w01_ncu.log 2>&1
ncu -i gpurun_out/w01_tile.ncu-rep --page raw --csv > gpurun_out/w01_raw.csv 2>/dev/null
ncu -i gpurun_out/w01_tile.ncu-rep --page source --csv > gpurun_out/w01_src.csv 2>/dev/null
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 3 python tools/sanitize_run.py 150 > gpurun_out/w01_racecheck.log 2>&1; echo "racecheck rc $?"; tail -2 gpurun_out/w01_racecheck.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 3 python tools/sanitize_run.py 300 > gpurun_out/w01_memcheck.log 2>&1; echo "memcheck rc $?"; tail -2 gpurun_out/w01_memcheck.log
