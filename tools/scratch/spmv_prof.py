import os, sys, numpy as np
sys.path.insert(0, os.getcwd())
from mac_b200 import _lib
n, m, band = 4_000_000, 40_000_000, 2000
rng = np.random.default_rng(0)
fi = np.arange(n - 1, dtype=np.int32)
a = rng.integers(0, n, size=m, dtype=np.int64)
b = a + rng.integers(2, band, size=m, dtype=np.int64); b = np.where(b >= n, a - (b - a), b)
ok = np.abs(a - b) > 1
ci, cj = a[ok].astype(np.int32), b[ok].astype(np.int32)
h = _lib.Handle(n, fi, fi + 1, np.ones(n - 1), ci, cj, np.ones(len(ci)))
h.set_x(np.ones(len(ci)))
h.spmv_engine(int(os.environ.get("ENGINE", "1")))
print(h.spmv_bench(5, False))
