python tools/scratch/rrstat2.py 2>&1 | tail -2
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python tools/scratch/gap.py 2>&1 | tail -3
