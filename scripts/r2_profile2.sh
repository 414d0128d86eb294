#!/bin/bash
# round-2: GPU tests + one full ncu capture of k_env_step inside the bench command (the .so of this snapshot: keep a copy next to the report)
cd "$GRAFT_REPO_ROOT"
export UHC_BENCH_SKIP_CPU=1
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -3
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_env_step -s 8 -c 1 -o gpurun_out/prof_env_r02b -f python bench.py --steps 8 --warmup 3 > gpurun_out/r02b_ncu.log 2>&1
tail -2 gpurun_out/r02b_ncu.log
cp uhc_b200/libuhc_b200.so gpurun_out/prof_env_r02b.so
