set -x
timeout 900 python -m pytest tests -m gpu -x -q --durations=8 -k "lanczos or fiedler or pure or fused or g2o or headline or er10k or zero_cand or petersen or solve_api" > gpurun_out/r2_tests2.log 2>&1; echo "pytest rc=$?"
tail -25 gpurun_out/r2_tests2.log
MACB_LIB=mac_b200/libmacb200_timing.so timeout 300 python tools/ptiming_pipe.py dense > gpurun_out/r2_ptiming_pipe.txt 2>&1; cat gpurun_out/r2_ptiming_pipe.txt
timeout 400 python bench.py --steps 20 --warmup 3 > gpurun_out/r2_bench2.json 2> gpurun_out/r2_bench2.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench2.json'))
print({k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d['roofline']['us_per_lanczos_step'], d['roofline']['share_of_timed_region'], d['config']['lanczos_steps_per_solve'], d.get('parity_check'))
PY
tail -3 gpurun_out/r2_bench2.err
MACB_NO_PIPE=1 timeout 400 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench2_nopipe.json 2>/dev/null
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench2_nopipe.json'))
print("nopipe", {k:d[k] for k in ('value','ms_per_step')}, d['roofline']['us_per_lanczos_step'])
PY
