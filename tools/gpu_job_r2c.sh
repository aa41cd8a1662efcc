#!/bin/bash
set -x
python tools/profile_merge.py
timeout 900 ncu --set full --clock-control none --kernel-name-base function -k 'regex:^(normals|intersect|merge_|segment_shift|fps_large|nn_sqdist|chamfer|pose_apply|edge_features|verifier_|attn_varlen|split_bf16|layernorm|gemm_bf16_tc)' -o gpurun_out/r2_full_merge python tools/profile_merge.py > gpurun_out/r2_ncu5.log 2>&1
ncu -i gpurun_out/r2_full_merge.ncu-rep --page raw --csv > gpurun_out/r2_full_merge_raw.csv 2>/dev/null
ls -la gpurun_out/r2_full_merge.ncu-rep gpurun_out/r2_full_merge_raw.csv; rm -f gpurun_out/r2_full_merge.ncu-rep
tail -2 gpurun_out/r2_ncu5.log
python -m pytest tests/test_gpu_reference_speed.py -x -q -m gpu -s 2>&1 | tail -8
python bench.py --steps 12 --warmup 3 > gpurun_out/r2_bench5.json 2> gpurun_out/r2_bench5.err; cut -c 1-300 gpurun_out/r2_bench5.json; tail -3 gpurun_out/r2_bench5.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_ref_arm.json 2>/dev/null; cut -c 1-400 gpurun_out/r2_ref_arm.json
