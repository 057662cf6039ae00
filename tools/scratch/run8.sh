timeout 600 python -m pytest tests -m gpu -x -q -k "farm" 2>&1 | tail -4
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r2_bench6.json 2> gpurun_out/r2_bench6.err; echo "bench rc=$?"; tail -5 gpurun_out/r2_bench6.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench6.json'))
print({k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d['config']['ksweep'])
r=d['roofline']; print({k:r[k] for k in ('achieved','frac','us_per_lanczos_step','share_of_timed_region','algorithmic_bytes_per_step','l2','hbm_spmv','standalone_spmv')})
print(d.get('parity_check'), d.get('cpu_baseline'))
PY
