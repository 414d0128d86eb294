#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_nn.py tests/test_gpu_ppo_c.py -x -q > gpurun_out/yt_tests.log 2>&1
tail -4 gpurun_out/yt_tests.log
timeout 300 python scripts/gemm_rates.py > gpurun_out/yt_gemm_rates.log 2>&1; cat gpurun_out/yt_gemm_rates.log
