#!/bin/bash
# third session of round 2: time keeper on all warps of its block + one more tick of lookahead: parity of the asynchronous forms, sweep, C3 / C4 lines
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "asynchronous or free_running" > $O/r2s3h_tests.log 2>&1; echo "tests rc=$?"; tail -3 $O/r2s3h_tests.log
timeout 600 python - > $O/r2s3h_forms.log 2>&1 <<'PY'
import sys
sys.path.insert(0, "tools")
sys.argv = ["x", "none"]
import async_forms as f
f.run("C3", "w4", (2956, 3552, 4144, 4732))
f.run("C4", "w4", (3548,))
PY
echo "forms rc=$?"; cat $O/r2s3h_forms.log
timeout 600 python bench.py --workload C3 --steps 3 --warmup 3 --no-extra --no-cpu > $O/r2s3h_bench_c3.json 2> $O/r2s3h_bench_c3.err; echo "c3 rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2s3h_bench_c3.json").read().strip().splitlines()[-1])
print(d["restarts_per_s"], d["stage_ms"], d["instances"], d["best_fval"], d.get("parity", {}).get("ok"))
PY
