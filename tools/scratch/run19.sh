python bench.py --steps 20 --warmup 3 --no-ksweep --no-hbm-spmv 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); c=d['config']; print(d['value'], d['ms_per_step'], d['e2e']['value'], c['lanczos_us_per_step'], c['lanczos_ms_per_iter'], c['other_kernels_ms_per_iter'], c['lanczos_steps_per_solve'])"
