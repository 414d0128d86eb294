#!/bin/bash
# round 2 experiment: latency-vs-residency curve of k_env_step (one wave = 2072 envs at 14 warps/SM) and register-cap cost
cd "$GRAFT_REPO_ROOT"
for lib in cta2 nosync reg112 reg96; do
  for E in 1036 2072 4096 8192; do
    echo -n "lib=$lib "; UHC_B200_SO=$PWD/build_variants/lib_$lib.so timeout 300 python scripts/quick_time.py $E 20 2>&1 | tail -1
  done
done > gpurun_out/r2_sweep.txt 2>&1
cat gpurun_out/r2_sweep.txt
