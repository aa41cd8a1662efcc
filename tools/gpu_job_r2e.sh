#!/bin/bash
set -x
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_res_ln --launch-skip 5 --launch-count 1 -o gpurun_out/r2e_resln python tools/bench_res_ln.py 9675 > gpurun_out/r2e_ncu.log 2>&1
tail -3 gpurun_out/r2e_ncu.log
ncu -i gpurun_out/r2e_resln.ncu-rep --page raw --csv > gpurun_out/r2e_resln_raw.csv 2>/dev/null
ncu -i gpurun_out/r2e_resln.ncu-rep --page source --csv --print-source sass > gpurun_out/r2e_resln_source.csv 2>/dev/null
ncu -i gpurun_out/r2e_resln.ncu-rep --page details > gpurun_out/r2e_resln_details.txt 2>/dev/null
ls -la gpurun_out/
