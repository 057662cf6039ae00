timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -2
python bench.py --steps 20 --warmup 5 --no-hbm-spmv 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); c=d['config']; print(d['value'], d['e2e']['value'], c['lanczos_us_per_step'], c['lanczos_steps_per_solve'], c['device_rr_fallbacks'], d['parity_check']['max_rel_err'], {k:(v['seconds'],v['max_rel_dlambda2_vs_reference']) for k,v in c['ksweep'].items() if isinstance(v,dict)})"
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
