timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize.py > gpurun_out/r2_sanitizer.txt 2>&1; echo "rc=$?"; tail -12 gpurun_out/r2_sanitizer.txt
