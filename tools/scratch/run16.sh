MACB_LIB=mac_b200/libmacb200_timing.so timeout 120 python tools/ptiming_pipe.py dense 2>&1 | tail -10
