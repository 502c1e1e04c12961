#!/bin/bash
# third session of round 2: free-running one-warp instances -- parity of every asynchronous form, then the C3 / C4 lines and the default bench
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "asynchronous or free_running" > $O/r2s3j_tests.log 2>&1; echo "tests rc=$?"; tail -3 $O/r2s3j_tests.log
timeout 600 python bench.py --workload C3 --steps 3 --warmup 3 --no-extra --no-cpu > $O/r2s3j_bench_c3.json 2> $O/r2s3j_bench_c3.err; echo "c3 rc=$?"
timeout 600 python bench.py --workload C4 --steps 3 --warmup 3 --no-extra --no-cpu > $O/r2s3j_bench_c4.json 2> $O/r2s3j_bench_c4.err; echo "c4 rc=$?"
( time timeout 900 python bench.py --steps 5 --warmup 3 > $O/r2s3j_bench.json 2> $O/r2s3j_bench.err ) 2> $O/r2s3j_time.txt; echo "bench rc=$?"; tail -3 $O/r2s3j_time.txt; tail -2 $O/r2s3j_bench.err
python - <<'PY'
import json
for f in ("_c3", "_c4", ""):
    d = json.loads(open(f"gpurun_out/r2s3j_bench{f}.json").read().strip().splitlines()[-1])
    print(f or "c2", d["restarts_per_s"], d["stage_ms"], d["instances"], d["best_fval"], d.get("rounds"), d.get("four_warp_blocks"), d.get("free_running"), (d.get("parity") or {}).get("ok"))
PY
