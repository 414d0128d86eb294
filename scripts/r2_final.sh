#!/bin/bash
# round-end dry run of what the driver does: smoke, GPU tests, both bench arms, the train workload
cd "$GRAFT_REPO_ROOT"
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -2
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_final_ref.json 2> gpurun_out/r2_final_ref.err; tail -c 700 gpurun_out/r2_final_ref.json; echo
timeout 900 python bench.py > gpurun_out/r2_final_bench.json 2> gpurun_out/r2_final_bench.err; python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2_final_bench.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches", "steps", "warmup")}, d["e2e"]["value"], d["roofline"]["frac"], d["roofline"]["traffic"], d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"], d["clocks"])
PY
timeout 900 python bench.py --workload train --steps 4 --warmup 3 > gpurun_out/r2_final_train.json 2> gpurun_out/r2_final_train.err; python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2_final_train.json").read().strip().splitlines()[-1])
print(d["value"], d["phases"]["sample_ms"], d["phases"]["update_ms"], d["gpu_launches"], d["roofline"]["frac"])
PY
