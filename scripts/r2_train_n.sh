#!/bin/bash
# train workload only on N GPUs of one box (the rollout scaling is the driver's own run)
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
N=${1:-8}
export UHC_BENCH_SKIP_CPU=1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --workload train --steps 3 --warmup 3 > gpurun_out/r2_train_n$N.json 2> gpurun_out/r2_train_n$N.err
python - <<PY
import json
d = json.loads(open("gpurun_out/r2_train_n$N.json").read().strip().splitlines()[-1])
print(d["value"], d["n_gpus"], d["phases"], d["replicas_identical"])
PY
tail -2 gpurun_out/r2_train_n$N.err
