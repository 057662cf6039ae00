ENGINE=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_spmv_jds -s 3 -c 1 -o gpurun_out/r2_spmv_jds python tools/scratch/spmv_prof.py 2>&1 | tail -3
ENGINE=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_spmv -s 3 -c 1 -o gpurun_out/r2_spmv_csr python tools/scratch/spmv_prof.py 2>&1 | tail -3
