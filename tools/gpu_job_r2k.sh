#!/bin/bash
set -x
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_tc3 --launch-skip 2 --launch-count 1 -o gpurun_out/r2k_tc3 python tools/bench_one_gemm.py 9675 4096 512 4 1 0 > gpurun_out/r2k_ncu.log 2>&1
tail -2 gpurun_out/r2k_ncu.log
ncu -i gpurun_out/r2k_tc3.ncu-rep --page source --csv --print-source sass > gpurun_out/r2k_tc3_source.csv 2>/dev/null
ncu -i gpurun_out/r2k_tc3.ncu-rep --page details > gpurun_out/r2k_tc3_details.txt 2>/dev/null
rm -f gpurun_out/r2k_tc3.ncu-rep
