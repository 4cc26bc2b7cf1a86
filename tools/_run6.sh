mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/t01_tests.log 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_stroke_tile -s 1 -c 1 -o gpurun_out/t01_tile -f python tools/run_c3.py 200000 2 > gpurun_out/t01_ncu.log 2>&1
ncu -i gpurun_out/t01_tile.ncu-rep --page raw --csv > gpurun_out/t01_raw.csv 2>/dev/null
ncu -i gpurun_out/t01_tile.ncu-rep --page source --csv > gpurun_out/t01_src.csv 2>/dev/null
tail -3 gpurun_out/t01_tests.log
