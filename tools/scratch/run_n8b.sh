timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2d_bench_n1_k20.json 2>/dev/null
for N in 2 4 8; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2962$N bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2d_bench_n$N.json 2> gpurun_out/r2d_bench_n$N.err; echo "bench N=$N rc=$?"; tail -2 gpurun_out/r2d_bench_n$N.err
done
python - <<PY
import json
for N in ('1_k20',2,4,8):
    d=json.loads(open(f'gpurun_out/r2d_bench_n{N}.json').read().strip().splitlines()[-1])
    print(N, {k:d[k] for k in ('value','ms_per_step','n_gpus')}, d['e2e']['value'], d['roofline']['us_per_lanczos_step'], d['roofline']['share_of_timed_region'], d['config']['lanczos_steps_per_solve'], {k:v['seconds'] for k,v in d['config']['ksweep'].items() if isinstance(v,dict)})
PY
timeout 600 python -m pytest tests -m gpu -q -k "farm or nccl or sweep" 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
