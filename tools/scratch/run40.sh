S=$(date +%s); python bench.py > gpurun_out/r2h_default.json 2> gpurun_out/r2h_default.err; echo rc=$? elapsed $(( $(date +%s) - S )) s
python - <<PY
import json
d=json.loads(open('gpurun_out/r2h_default.json').read().strip().splitlines()[-1])
print(d['value'], d['steps'], d['e2e']['value'], d['config']['lanczos_steps_per_solve'], d['roofline']['frac'], d['roofline']['hbm_spmv']['frac'], d['cpu_baseline']['value'], d['parity_check']['max_rel_err'], d['gpu_launches'])
PY
