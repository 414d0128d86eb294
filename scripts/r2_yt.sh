#!/bin/bash
# transposed-activation epilogue: tests, per-GEMM rates, train bench A/B
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_nn.py tests/test_gpu_ppo_c.py tests/test_gpu_agent.py -x -q > gpurun_out/yt_tests.log 2>&1
tail -5 gpurun_out/yt_tests.log
timeout 300 python scripts/gemm_rates.py > gpurun_out/yt_gemm_rates.log 2>&1; tail -30 gpurun_out/yt_gemm_rates.log
timeout 600 python bench.py --workload train --steps 4 --warmup 3 > gpurun_out/yt_train.json 2> gpurun_out/yt_train.err; cat gpurun_out/yt_train.json
UHC_TC_TMA_STORE=0 timeout 600 python bench.py --workload train --steps 4 --warmup 3 > gpurun_out/yt_train_legacy.json 2>> gpurun_out/yt_train.err; cat gpurun_out/yt_train_legacy.json
