#!/bin/bash
# round-2: full ncu capture of k_env_step inside the bench command + launch list of one bench run (the .so of this snapshot is copied next to the report)
cd "$GRAFT_REPO_ROOT"
export UHC_BENCH_SKIP_CPU=1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_env_step -s 8 -c 1 -o gpurun_out/prof_env_r02c -f python bench.py --steps 8 --warmup 3 > gpurun_out/r02c_ncu.log 2>&1
tail -2 gpurun_out/r02c_ncu.log
cp uhc_b200/libuhc_b200.so gpurun_out/prof_env_r02c.so
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 300 --csv --log-file gpurun_out/r02c_bench_launches.csv python bench.py --steps 8 --warmup 3 > gpurun_out/r02c_ncu_a.log 2>&1
tail -1 gpurun_out/r02c_ncu_a.log | cut -c1-200
