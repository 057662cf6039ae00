"""mac_b200: B200-native Frank-Wolfe / Fiedler hot path behind the MAC API.

Drop-in surface (mirrors the reference's `mac` package for this path):
    mac_b200.solvers.MAC                       (mac/solvers/mac.py:15-225)
    mac_b200.optimization.frankwolfe.frank_wolfe
    mac_b200.optimization.constraints.solve_subset_box_lp
    mac_b200.utils.{graphs,fiedler,rounding,conversions}
All numerics run in hand-written sm_100a CUDA kernels behind the C-ABI in
`include/macb200.h` (libmacb200.so, loaded with ctypes).  There is no CPU fallback.
"""
__version__ = "0.1.0"
