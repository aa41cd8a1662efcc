#!/bin/bash
# round 2, session 2: evidence for the fused residual+LayerNorm projection and the warp-MMA local attention
set -x
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "gemm_res_ln or attention_local" > gpurun_out/r2s2_sanitizer_memcheck.log 2>&1; echo "sanitizer rc=$?"; tail -4 gpurun_out/r2s2_sanitizer_memcheck.log
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2s2_launches_T100.csv python tools/profile_step.py --steps 100 --iters 2 --chamfer > gpurun_out/r2s2_ncu1.log 2>&1
tail -1 gpurun_out/r2s2_ncu1.log | cut -c 1-200
K='regex:^(gemm_res_ln|attention_local|attention_ws|gemm_bf16_tc3|layernorm|sa_fused|sa_resident)'
timeout 900 ncu --set full --clock-control none --kernel-name-base function -k "$K" -o gpurun_out/r2s2_full python tools/profile_step.py --steps 1 --iters 1 > gpurun_out/r2s2_ncu2.log 2>&1
ncu -i gpurun_out/r2s2_full.ncu-rep --page raw --csv > gpurun_out/r2s2_full_raw.csv 2>/dev/null
ls -la gpurun_out/r2s2_full.ncu-rep gpurun_out/r2s2_full_raw.csv; rm -f gpurun_out/r2s2_full.ncu-rep
tail -2 gpurun_out/r2s2_ncu2.log
