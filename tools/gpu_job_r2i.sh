#!/bin/bash
set -x
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attention_ws --launch-skip 3 --launch-count 1 -o gpurun_out/r2i_attn python tools/bench_attention.py 32 4 > gpurun_out/r2i_ncu.log 2>&1
tail -3 gpurun_out/r2i_ncu.log
ncu -i gpurun_out/r2i_attn.ncu-rep --page source --csv --print-source sass > gpurun_out/r2i_attn_source.csv 2>/dev/null
ncu -i gpurun_out/r2i_attn.ncu-rep --page details > gpurun_out/r2i_attn_details.txt 2>/dev/null
rm -f gpurun_out/r2i_attn.ncu-rep
