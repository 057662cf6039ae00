MACB_LIB=mac_b200/libmacb200_timing.so timeout 120 python tools/ptiming_pipe.py dense 2>&1 | tail -10
python bench.py --steps 20 --warmup 5 --no-ksweep --no-hbm-spmv 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); c=d['config']; print(d['value'], d['ms_per_step'], d['e2e']['value'], c['lanczos_us_per_step'])"
