#!/bin/bash
# fused epilogues: tests, train bench A/B
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/yt_tests.log 2>&1
tail -15 gpurun_out/yt_tests.log
timeout 600 python bench.py --workload train --steps 4 --warmup 3 > gpurun_out/yt_train.json 2> gpurun_out/yt_train.err; cat gpurun_out/yt_train.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['phases'], d['gpu_launches'])"
