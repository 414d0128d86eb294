#!/bin/bash
# GPU tests, kernel-only timing and the bench line of the current tree
cd "$GRAFT_REPO_ROOT"
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -3
timeout 300 python scripts/quick_time.py 4096 20
UHC_BENCH_SKIP_CPU=1 timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/r2_bench_c.json 2> gpurun_out/r2_bench_c.err; python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2_bench_c.json").read().strip().splitlines()[-1])
print("bench value %.0f e2e %.0f kernel_ms %.3f ms/step %.3f counters %s" % (d["value"], d["e2e"]["value"], d["roofline"]["kernel_ms"], d["ms_per_step"], d["env_counters"]))
PY
tail -3 gpurun_out/r2_bench_c.err
