timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-hbm-spmv > gpurun_out/r2_bench7.json 2> gpurun_out/r2_bench7.err; echo "bench rc=$?"; tail -3 gpurun_out/r2_bench7.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench7.json'))
print({k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d['roofline']['us_per_lanczos_step'], d['config']['lanczos_steps_per_solve'], d['gpu_launches'], {k:v['seconds'] for k,v in d['config']['ksweep'].items() if isinstance(v,dict)})
PY
