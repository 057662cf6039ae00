for lib in libmacb200_oldgap.so libmacb200.so; do echo "== $lib"; MACB_LIB=mac_b200/$lib python tools/scratch/gap.py 2>&1 | tail -3; done
