timeout 600 python -m pytest tests -m gpu -x -q -k "spmv" 2>&1 | tail -3
timeout 600 python tools/big_spmv.py 2>&1 | cut -c1-330
