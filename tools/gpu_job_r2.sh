#!/bin/bash
# one GPU-box job: tests, config-3 dynamics, bench, ncu launch list, ncu --set full of every library kernel
set -x
python -m pytest tests/test_gpu_engine.py tests/test_gpu_dropin.py -x -q -m gpu 2>&1 | tail -8 > gpurun_out/r2_t7.txt; cat gpurun_out/r2_t7.txt
python tools/probe_config3.py --repeat 2 --accept-bias -1.0 2>&1 | grep -v "^objects" > gpurun_out/r2_probe3c.txt; cat gpurun_out/r2_probe3c.txt
python bench.py --no-cpu-baseline > gpurun_out/r2_bench3.json 2> gpurun_out/r2_bench3.err; cut -c 1-300 gpurun_out/r2_bench3.json; tail -3 gpurun_out/r2_bench3.err
K='regex:^(fps_|ball_query|sa_|gemm_|attention_|attn_|layernorm|vq_|embed|combine|mean_pool|ddpm|step_|pose_apply|edge_features|verifier_|normals|intersect|merge_|segment_shift|scatter_rows|nn_sqdist|local_segments|split_bf16|group_)'
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches.csv python tools/profile_step.py --steps 4 --iters 2 > gpurun_out/r2_ncu1.log 2>&1
timeout 1500 ncu --set full --clock-control none --import-source on --kernel-name-base function -k "$K" -o gpurun_out/r2_full python tools/profile_step.py --steps 1 --iters 2 > gpurun_out/r2_ncu2.log 2>&1
ncu -i gpurun_out/r2_full.ncu-rep --page raw --csv > gpurun_out/r2_full_raw.csv 2>/dev/null
ls -la gpurun_out/r2_full.ncu-rep gpurun_out/r2_full_raw.csv
SZ=$(stat -c %s gpurun_out/r2_full.ncu-rep 2>/dev/null || echo 0); if [ "$SZ" -gt 45000000 ]; then rm gpurun_out/r2_full.ncu-rep; fi
tail -3 gpurun_out/r2_ncu2.log
