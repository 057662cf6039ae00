"""L2 -> SM read bandwidth at a few buffer sizes (macb_measure_l2_bandwidth).  Scratch/measurement tool."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mac_b200 import _lib
out = {}
for mb in (4, 8, 16, 24, 32, 48, 64, 96):
    out[f"{mb}MB"] = round(_lib.measure_l2_bandwidth(-1, mb << 20, max(4, 256 // mb)), 1)
print(json.dumps({"l2_read_GBs": out}))
