for g in city10000 sphere2500; do MACB_LIB=mac_b200/libmacb200_timing.so timeout 120 python tools/ptiming_pipe.py $g 2>&1 | tail -11; done
