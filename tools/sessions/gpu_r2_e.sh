#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > $O/r2e_pytest.log 2>&1; echo "pytest rc=$?" > $O/r2e_status.txt
for w in C2 C3 C4; do
  timeout 300 python bench.py --workload $w --steps 3 --warmup 3 --no-cpu > $O/r2e_bench_${w}.json 2> $O/r2e_bench_${w}.err; echo "bench $w rc=$?" >> $O/r2e_status.txt
done
timeout 300 python bench.py --workload C3 --round-size 888 --steps 3 --warmup 3 --no-cpu > $O/r2e_bench_C3_r888.json 2> $O/r2e_bench_C3_r888.err
tail -3 $O/r2e_pytest.log; cat $O/r2e_status.txt
for f in $O/r2e_bench_*.json; do python - "$f" <<'PY'
import json,sys
d=json.load(open(sys.argv[1])); print(sys.argv[1], 'restarts/s %.0f'%d['restarts_per_s'], 'local ms %.3f'%d['stage_ms']['local'], 'launches', d['gpu_launches'], 'best', d['best_fval'])
PY
done
