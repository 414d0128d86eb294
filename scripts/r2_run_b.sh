#!/bin/bash
cd "$GRAFT_REPO_ROOT"
timeout 900 python -m pytest tests/test_gpu_ppo_c.py tests/test_gpu_agent.py tests/test_gpu_dropin.py -q -m gpu -x 2>&1 | tail -15
UHC_BENCH_SKIP_CPU=1 timeout 600 python bench.py --workload train --steps 4 --warmup 3 > gpurun_out/r2_train_c.json 2> gpurun_out/r2_train_c.err; tail -c 1500 gpurun_out/r2_train_c.json; tail -3 gpurun_out/r2_train_c.err
