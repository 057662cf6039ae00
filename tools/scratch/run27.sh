for tool in racecheck synccheck; do
timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 python tools/scratch/san2.py > gpurun_out/r2_sanitizer_$tool.txt 2>&1; echo "$tool rc=$?"; tail -6 gpurun_out/r2_sanitizer_$tool.txt
done
