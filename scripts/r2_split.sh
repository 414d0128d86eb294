#!/bin/bash
# alignment-group size of the substep barrier on the round-2 kernel: the libraries named by the arguments (build_variants/lib_<name>.so), "default" = the in-tree one
cd "$GRAFT_REPO_ROOT"
for v in "$@"; do
  if [ $v = default ]; then unset UHC_B200_SO; else export UHC_B200_SO=$PWD/build_variants/lib_$v.so; fi
  echo -n "$v: "; timeout 300 python scripts/quick_time.py 4096 20 2>&1 | tail -1
  UHC_BENCH_SKIP_CPU=1 timeout 300 python bench.py --steps 20 --warmup 3 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('   bench %.0f kernel_ms %.3f' % (d['value'], d['roofline']['kernel_ms']))"
done
