timeout 900 python -m pytest tests -m gpu -x -q -k "farm" 2>&1 | tail -3
for N in 1 2; do
if [ $N = 1 ]; then python bench.py --steps 20 --warmup 5 --no-hbm-spmv > gpurun_out/r2e_n$N.json 2>gpurun_out/r2e_n$N.err; else
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2963$N bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2e_n$N.json 2> gpurun_out/r2e_n$N.err; fi
python - <<PY
import json
d=json.loads(open('gpurun_out/r2e_n$N.json').read().strip().splitlines()[-1])
print($N, d['value'], d['e2e']['value'], {k:(v['seconds'], v['streams_per_gpu'], v['max_rel_dlambda2_vs_reference']) for k,v in d['config']['ksweep'].items() if isinstance(v,dict)})
PY
done
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
