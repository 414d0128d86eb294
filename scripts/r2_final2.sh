#!/bin/bash
# round-end dry run: smoke, GPU tests, rollout + train bench, per-GEMM rates, ncu captures of the pair kernel, launch list of the default bench
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/f2_bench.json 2> gpurun_out/f2_bench.err; python - <<'PY'
import json
d = json.loads(open("gpurun_out/f2_bench.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches", "steps", "warmup")}, d["e2e"]["value"], d["roofline"]["frac"], d["roofline"]["traffic"], d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"], d["clocks"])
PY
timeout 900 python bench.py --workload train --steps 4 --warmup 3 > gpurun_out/f2_train.json 2> gpurun_out/f2_train.err; python - <<'PY'
import json
d = json.loads(open("gpurun_out/f2_train.json").read().strip().splitlines()[-1])
print(d["value"], d["phases"]["sample_ms"], d["phases"]["update_ms"], d["gpu_launches"], d["roofline"]["frac"])
PY
timeout 300 python scripts/gemm_rates.py > gpurun_out/f2_gemm_rates.log 2>&1; cat gpurun_out/f2_gemm_rates.log
for cfg in "fwd 657 2048" "fwd 2048 1024" "dact 1024 2048" "dw 2048 1024"; do
  tag=$(echo $cfg | tr ' ' '_')
  timeout 300 ncu --set full --import-source on --clock-control none -k regex:k_linear_tc2 -s 2 -c 1 -f -o gpurun_out/f2_$tag python scripts/gemm_one.py $cfg > gpurun_out/f2_ncu_$tag.log 2>&1
  echo "$cfg rc=$?"
done
UHC_BENCH_SKIP_CPU=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/f2_bench_launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/f2_ncu_bench.log 2>&1; echo "launch list rc=$?"
