timeout 600 python -m pytest tests -m gpu -x -q -k "farm" 2>&1 | tail -2
for i in 1 2 3 4; do python - <<'PY'
import os, sys, json
sys.path.insert(0, os.getcwd())
import bench
out = bench.ksweep(1, 0, 0)
print({k: (round(v["seconds"], 4), v["streams_per_gpu"]) for k, v in out.items() if isinstance(v, dict)})
PY
done
