#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 100 python -m pytest tests/test_gpu_env.py -q -x -k "v5_v6 or fp32_kernels_within" 2>&1 | tail -3
UHC_BENCH_SKIP_CPU=1 timeout 60 python bench.py --steps 20 --warmup 3 2>/dev/null > gpurun_out/last_bench.json; python -c "
import json; d=json.loads(open('gpurun_out/last_bench.json').read().strip().splitlines()[-1]); print('rollout %.0f e2e %.0f kernel_ms %.3f' % (d['value'], d['e2e']['value'], d['roofline']['kernel_ms']))"
