for N in 8 4; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2961$N bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err; echo "bench N=$N rc=$?"; tail -2 gpurun_out/r2_bench_n$N.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r2_bench_n$N.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, d['e2e']['value'], d['roofline']['us_per_lanczos_step'], d['roofline']['share_of_timed_region'], d['config']['lanczos_steps_per_solve'], {k:v['seconds'] for k,v in d['config']['ksweep'].items() if isinstance(v,dict)})
PY
done
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-hbm-spmv > gpurun_out/r2_bench_n1b.json 2>/dev/null
python - <<PY
import json
d=json.loads(open('gpurun_out/r2_bench_n1b.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, d['e2e']['value'], d['roofline']['us_per_lanczos_step'], d['roofline']['share_of_timed_region'], d['config']['lanczos_steps_per_solve'])
PY
