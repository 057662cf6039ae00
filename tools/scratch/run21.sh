timeout 900 python -m pytest tests -m gpu -x -q -k "dense or topk or round" 2>&1 | tail -3
python - <<'PY'
import time, numpy as np
from mac_b200.optimization.constraints import solve_subset_box_lp
rng=np.random.default_rng(0); g=rng.random(1_000_000)
for i in range(4):
    t=time.perf_counter(); s=solve_subset_box_lp(g, 200000); print("solve_subset_box_lp m=1M call", i, "%.2f ms" % ((time.perf_counter()-t)*1e3))
PY
