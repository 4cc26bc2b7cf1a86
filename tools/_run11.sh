mkdir -p gpurun_out
timeout 80 compute-sanitizer --tool memcheck --error-exitcode 3 python tools/sanitize_run.py 300 > gpurun_out/z01_memcheck.log 2>&1; echo "memcheck rc $?"; tail -2 gpurun_out/z01_memcheck.log
