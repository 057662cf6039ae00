MACB_LIB=mac_b200/libmacb200_timing.so timeout 300 python tools/ptiming_pipe.py dense 2>&1 | tail -13
timeout 900 python -m pytest tests -m gpu -x -q -k "lanczos or fiedler or pure or fused or g2o or headline or er10k or zero_cand or petersen or solve_api" 2>&1 | tail -4
timeout 400 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench3.json 2> gpurun_out/r2_bench3.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench3.json'))
print({k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d['roofline']['us_per_lanczos_step'], d['roofline']['share_of_timed_region'], d['config']['lanczos_steps_per_solve'])
PY
