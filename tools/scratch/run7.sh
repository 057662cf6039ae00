timeout 400 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench5.json 2> gpurun_out/r2_bench5.err; echo "bench rc=$?"; tail -3 gpurun_out/r2_bench5.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench5.json'))
print({k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d['roofline']['us_per_lanczos_step'], d['roofline']['share_of_timed_region'], d['config']['lanczos_steps_per_solve'], d['gpu_launches'])
PY
MACB_HOST_RR=1 timeout 400 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench5h.json 2>/dev/null
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench5h.json'))
print("host RR", {k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d['roofline']['us_per_lanczos_step'], d['roofline']['share_of_timed_region'], d['config']['lanczos_steps_per_solve'], d['gpu_launches'])
PY
