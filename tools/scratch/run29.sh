timeout 600 python -m pytest tests -m gpu -x -q -k "headline or engines or er10k or golden" 2>&1 | tail -2
python bench.py --steps 20 --warmup 5 --no-ksweep --no-hbm-spmv 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); c=d['config']; print(d['value'], d['ms_per_step'], d['e2e']['value'], c['lanczos_us_per_step'])"
timeout 400 compute-sanitizer --tool racecheck python tools/scratch/san3.py > gpurun_out/r2_racecheck2.txt 2>&1; echo rc=$?; grep -v "Host Frame\|^=========         at\|^=========     at" gpurun_out/r2_racecheck2.txt | tail -12
