#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -3
UHC_BENCH_SKIP_CPU=1 timeout 600 python bench.py --steps 20 --warmup 3 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('rollout %.0f e2e %.0f kernel_ms %.3f' % (d['value'], d['e2e']['value'], d['roofline']['kernel_ms']))"
