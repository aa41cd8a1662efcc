#!/bin/bash
set -x
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attention_local --launch-skip 3 --launch-count 1 -o gpurun_out/r2f_attn_local python tools/bench_attention.py 19 4 > gpurun_out/r2f_ncu.log 2>&1
tail -3 gpurun_out/r2f_ncu.log
ncu -i gpurun_out/r2f_attn_local.ncu-rep --page source --csv --print-source sass > gpurun_out/r2f_attn_local_source.csv 2>/dev/null
ncu -i gpurun_out/r2f_attn_local.ncu-rep --page details > gpurun_out/r2f_attn_local_details.txt 2>/dev/null
rm -f gpurun_out/r2f_attn_local.ncu-rep
