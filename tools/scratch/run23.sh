timeout 600 python -m pytest tests -m gpu -x -q -k "farm" 2>&1 | tail -3
for st in 1 3 6 9; do MACB_KSWEEP_STREAMS=$st python - <<'PY'
import os, sys, json
sys.path.insert(0, os.getcwd())
import bench
out = bench.ksweep(1, 0, 0)
print("streams", os.environ["MACB_KSWEEP_STREAMS"], {k: (round(v["seconds"], 4), v["max_rel_dlambda2_vs_reference"], v["selected_ok"]) for k, v in out.items() if isinstance(v, dict)})
PY
done
