set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
python tools/l2probe.py > gpurun_out/r2_l2probe.json 2> gpurun_out/r2_l2probe.err
timeout 1200 python -m pytest tests -m gpu -x -q --durations=15 > gpurun_out/r2_tests1.log 2>&1; echo "pytest rc=$?"
tail -30 gpurun_out/r2_tests1.log
timeout 400 python bench.py --steps 20 --warmup 3 > gpurun_out/r2_bench1.json 2> gpurun_out/r2_bench1.err; echo "bench rc=$?"
cat gpurun_out/r2_bench1.json; tail -3 gpurun_out/r2_bench1.err
cat gpurun_out/r2_l2probe.json
