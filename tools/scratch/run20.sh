timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3
python bench.py > gpurun_out/r2d_bench_n1.json 2> gpurun_out/r2d_bench_n1.err; tail -c 600 gpurun_out/r2d_bench_n1.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2d_bench_ref.json 2> gpurun_out/r2d_bench_ref.err; cat gpurun_out/r2d_bench_ref.json | cut -c1-600
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
