timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -2
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/final_n1.json 2> gpurun_out/final_n1.err; python - <<PY
import json
d=json.loads(open('gpurun_out/final_n1.json').read().strip().splitlines()[-1])
print(d['value'], d['e2e']['value'], d['config']['lanczos_us_per_step'], d['config']['lanczos_steps_per_solve'], d['roofline']['frac'], d['roofline']['hbm_spmv']['frac'], d['parity_check']['max_rel_err'], d['gpu_launches'], {k:v['seconds'] for k,v in d['config']['ksweep'].items() if isinstance(v,dict)})
PY
