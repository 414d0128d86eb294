#!/bin/bash
cd "$GRAFT_REPO_ROOT"
for v in default tc3; do
  if [ $v = default ]; then unset UHC_B200_SO; else export UHC_B200_SO=$PWD/build_variants/lib_$v.so; fi
  echo "== $v"
  UHC_BENCH_SKIP_CPU=1 timeout 300 python bench.py --steps 20 --warmup 3 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('   rollout %.0f ms/step %.3f kernel_ms %.3f' % (d['value'], d['ms_per_step'], d['roofline']['kernel_ms']))"
  UHC_BENCH_SKIP_CPU=1 timeout 600 python bench.py --workload train --steps 4 --warmup 3 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('   train %.0f sample %.1f update %.1f' % (d['value'], d['phases']['sample_ms'], d['phases']['update_ms']))"
done
