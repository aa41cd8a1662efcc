#!/bin/bash
# round 2, session 2: fused residual projection + LayerNorm
set -x
timeout 300 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "gemm_res_ln" 2>&1 | tail -15
timeout 120 python tools/bench_res_ln.py 2>&1 | tail -6
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -15
timeout 600 python bench.py --steps 8 --warmup 3 > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err; cut -c 1-400 gpurun_out/r2d_bench.json; tail -3 gpurun_out/r2d_bench.err
