for N in 8 4; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2965$N bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2f_bench_n$N.json 2> gpurun_out/r2f_bench_n$N.err; echo "bench N=$N rc=$?"; tail -2 gpurun_out/r2f_bench_n$N.err | cut -c1-300
python - <<PY
import json
d=json.loads(open('gpurun_out/r2f_bench_n$N.json').read().strip().splitlines()[-1])
print($N, d['value'], d['e2e']['value'], d['roofline']['us_per_lanczos_step'], {k:(round(v['seconds'],4), v['streams_per_gpu'], v['max_rel_dlambda2_vs_reference'], v['selected_ok']) for k,v in d['config']['ksweep'].items() if isinstance(v,dict)})
PY
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29659 bench.py --impl reference --gpus 8 --steps 2 --warmup 1 2>/dev/null | tail -1 | cut -c1-300
