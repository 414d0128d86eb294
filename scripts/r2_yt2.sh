#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/yt_tests.log 2>&1
tail -5 gpurun_out/yt_tests.log
