#!/bin/bash
# third session of round 2: all GPU tests, bench as the driver runs it, reference arm, smoke
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > $O/r2s3_tests.log 2>&1; echo "tests rc=$?"; tail -3 $O/r2s3_tests.log
( time timeout 900 python bench.py --steps 5 --warmup 3 > $O/r2s3_bench.json 2> $O/r2s3_bench.err ) 2> $O/r2s3_time.txt; echo "bench rc=$?"; tail -3 $O/r2s3_time.txt; tail -2 $O/r2s3_bench.err
python -c "import __graft_entry__ as g; g.smoke()" > $O/r2s3_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $O/r2s3_smoke.log
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2s3_bench.json").read().strip().splitlines()[-1])
print("value %.3g frac %.3f e2e %.3g" % (d["value"], d["roofline"]["frac"], d["e2e"]["value"]), d.get("parity", {}).get("ok"))
for k, v in (d.get("per_workload") or {}).items():
    print(k, {kk: v.get(kk) for kk in ("restarts_per_s", "local_ms", "instances", "instance_form", "best_fval", "rounds", "frac")})
PY
