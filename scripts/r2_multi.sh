#!/bin/bash
# round-2 multi-GPU evidence: train workload (gradient all-reduce inside the timed region) and the rollout headline on N GPUs of one box
cd "$GRAFT_REPO_ROOT"
N=${1:-2}
export UHC_BENCH_SKIP_CPU=1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --workload train --steps 4 --warmup 3 > gpurun_out/r2_train_n$N.json 2> gpurun_out/r2_train_n$N.err
tail -c 1800 gpurun_out/r2_train_n$N.json; tail -3 gpurun_out/r2_train_n$N.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/r2_rollout_n$N.json 2> gpurun_out/r2_rollout_n$N.err
tail -c 600 gpurun_out/r2_rollout_n$N.json; tail -3 gpurun_out/r2_rollout_n$N.err
