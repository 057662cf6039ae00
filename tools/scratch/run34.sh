python tools/scratch/sweepvar.py sphere2500 2>&1 | tail -12
