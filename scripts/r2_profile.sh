#!/bin/bash
# round-2 profiles of the bench command itself: launch list (shares), one full capture of k_env_step with the bench's L2 flush, GEMM capture
cd "$GRAFT_REPO_ROOT"
export UHC_BENCH_SKIP_CPU=1
timeout 600 python -m pytest tests -q -m gpu -x 2>&1 | tail -2
timeout 300 python scripts/quick_time.py 4096 20
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 300 --csv --log-file gpurun_out/r02_bench_launches.csv python bench.py --steps 8 --warmup 3 > gpurun_out/r02_ncu_a.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_env_step -s 8 -c 2 -o gpurun_out/prof_env_r02 -f python bench.py --steps 8 --warmup 3 > gpurun_out/r02_ncu_b.log 2>&1
tail -3 gpurun_out/r02_ncu_a.log gpurun_out/r02_ncu_b.log
python bench.py --steps 20 --warmup 3 > gpurun_out/r2_bench_b.json 2> gpurun_out/r2_bench_b.err; tail -c 1500 gpurun_out/r2_bench_b.json
