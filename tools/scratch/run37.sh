for lib in libmacb200.so libmacb200_pf.so; do echo "== $lib"; MACB_LIB=mac_b200/$lib python tools/scratch/rrstat2.py 2>&1 | tail -2 | cut -c1-60
MACB_LIB=mac_b200/$lib python bench.py --steps 20 --warmup 5 --no-hbm-spmv --no-ksweep --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); c=d['config']; print(d['value'], c['lanczos_us_per_step'], c['lanczos_steps_per_solve'])"
MACB_LIB=mac_b200/$lib python tools/scratch/gap.py 2>&1 | tail -3 | cut -c1-120
done
