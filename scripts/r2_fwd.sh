#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
for m in 16384 4096 2048; do
  echo "== UHC_TC_PAIR_MINM=$m"; UHC_TC_PAIR_MINM=$m timeout 120 python scripts/fwd_rates.py 4096
done > gpurun_out/fwd_rates.log 2>&1
cat gpurun_out/fwd_rates.log
