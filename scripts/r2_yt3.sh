#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 300 python scripts/gemm_rates.py > gpurun_out/yt_gemm_rates.log 2>&1; cat gpurun_out/yt_gemm_rates.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/yt_train_launches.csv python bench.py --workload train --steps 1 --warmup 1 > gpurun_out/yt_ncu_train.log 2>&1
python - <<'PY'
import csv, collections
rows = []
with open("gpurun_out/yt_train_launches.csv") as f:
    lines = [l for l in f if not l.startswith("==")]
r = csv.DictReader(lines)
agg = collections.defaultdict(lambda: [0, 0.0])
for row in r:
    try: v = float(row["Metric Value"].replace(",", ""))
    except Exception: continue
    unit = row.get("Metric Unit", "ns")
    if unit in ("us", "usecond"): v *= 1e3
    elif unit in ("ms", "msecond"): v *= 1e6
    k = row["Kernel Name"][:90]
    agg[k][0] += 1; agg[k][1] += v
tot = sum(v[1] for v in agg.values())
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:30]:
    print(f"{t/1e6:9.3f} ms {100*t/tot:5.1f}%  n={n:5d}  {k}")
print("total", tot / 1e6, "ms")
PY
