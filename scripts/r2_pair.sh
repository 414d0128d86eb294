#!/bin/bash
# CTA-pair kernel bring-up: the unit tests that reach it under a short timeout, then rates
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 150 python -m pytest tests/test_gpu_nn.py -x -q -k "transposed or fused_dx or gemm_shapes" > gpurun_out/pair_tests.log 2>&1
echo "rc=$?"; tail -25 gpurun_out/pair_tests.log
nvidia-smi --query-gpu=utilization.gpu,memory.used --format=csv
timeout 120 python scripts/gemm_rates.py > gpurun_out/pair_gemm_rates.log 2>&1; echo "rc=$?"; cat gpurun_out/pair_gemm_rates.log | tail -16
