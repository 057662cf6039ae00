for mb in 4 5 6 8; do MACB_SJ_MINB=$mb python - <<'PY'
import os, sys, numpy as np
sys.path.insert(0, os.getcwd())
exec(open('tools/scratch/spmv_prof.py').read().replace('print(h.spmv_bench(5, False))','ms,by=h.spmv_bench(20, False); print("minb", os.environ.get("MACB_SJ_MINB"), "ms %.4f" % ms, "GB/s %.0f" % (by/ms/1e6), "frac %.3f" % (by/ms/1e6/6454))'))
PY
done
