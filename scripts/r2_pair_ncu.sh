#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
for cfg in "dact 1024 2048" "fwd 657 2048"; do
  tag=$(echo $cfg | tr ' ' '_')
  timeout 300 ncu --set full --import-source on --clock-control none -k regex:k_linear_tc2 -s 2 -c 1 -f -o gpurun_out/pair3_$tag python scripts/gemm_one.py $cfg > gpurun_out/pair3_ncu_$tag.log 2>&1
  echo "$cfg rc=$?"
done
