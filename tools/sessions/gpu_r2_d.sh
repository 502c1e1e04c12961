#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
TIKTAK_NM_MODE=324 timeout 900 python -m pytest tests -m gpu -x -q > $O/r2d_pytest_324.log 2>&1; echo "pytest 324 rc=$?" > $O/r2d_status.txt
for w in C2 C3 C4; do
  for md in 321 324; do
    TIKTAK_NM_MODE=$md timeout 300 python bench.py --workload $w --steps 3 --warmup 3 --no-cpu > $O/r2d_bench_${w}_m$md.json 2> $O/r2d_bench_${w}_m$md.err; echo "bench $w m$md rc=$?" >> $O/r2d_status.txt
  done
done
tail -3 $O/r2d_pytest_324.log; cat $O/r2d_status.txt
for f in $O/r2d_bench_*_m3*.json; do python - "$f" <<'PY'
import json,sys
d=json.load(open(sys.argv[1])); print(sys.argv[1], 'restarts/s %.0f'%d['restarts_per_s'], 'local ms %.3f'%d['stage_ms']['local'], 'launches', d['gpu_launches'], 'best', d['best_fval'])
PY
done
