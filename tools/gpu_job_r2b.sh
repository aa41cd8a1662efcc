#!/bin/bash
set -x
python -m pytest tests/test_gpu_kernels.py tests/test_gpu_engine.py -x -q -m gpu 2>&1 | tail -4
for s in 1 2 4; do echo "== config2 slots $s"; python bench.py --workload config2 --slots $s --no-cpu-baseline --steps 8 2>/dev/null | cut -c 1-230; done
python bench.py > gpurun_out/r2_bench4.json 2> gpurun_out/r2_bench4.err; cut -c 1-300 gpurun_out/r2_bench4.json; tail -3 gpurun_out/r2_bench4.err
# launch list of a run with merges (T = 100) and a full capture of the merge / verify / metric / Chamfer kernels
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_T100.csv python tools/profile_step.py --steps 100 --iters 2 --chamfer > gpurun_out/r2_ncu3.log 2>&1
tail -1 gpurun_out/r2_ncu3.log | cut -c 1-300
K='regex:^(normals|intersect|merge_|segment_shift|fps_large|nn_sqdist|chamfer|pose_apply|edge_features|verifier_|attn_varlen|split_bf16|gemm_bf16_tc_kernel|scatter_rows)'
timeout 1200 ncu --set full --clock-control none --kernel-name-base function -k "$K" -c 120 -o gpurun_out/r2_full_b python tools/profile_step.py --steps 100 --iters 2 --chamfer > gpurun_out/r2_ncu4.log 2>&1
ncu -i gpurun_out/r2_full_b.ncu-rep --page raw --csv > gpurun_out/r2_full_b_raw.csv 2>/dev/null
ls -la gpurun_out/r2_full_b.ncu-rep gpurun_out/r2_full_b_raw.csv; rm -f gpurun_out/r2_full_b.ncu-rep
tail -2 gpurun_out/r2_ncu4.log
