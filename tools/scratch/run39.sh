timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29671 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2g_n2.json 2> gpurun_out/r2g_n2.err; echo rc=$?
python - <<PY
import json
d=json.loads(open('gpurun_out/r2g_n2.json').read().strip().splitlines()[-1])
print(d['value'], d['e2e']['value'], d['config']['lanczos_steps_per_solve'], {k:(v['seconds'], v['streams_per_gpu'], v['selected_ok']) for k,v in d['config']['ksweep'].items() if isinstance(v,dict)})
PY
timeout 600 python -m pytest tests -m gpu -q -k "farm" 2>&1 | tail -2
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29672 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 2>/dev/null | tail -1 | cut -c1-200
