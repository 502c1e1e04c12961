#!/bin/bash
# third session of round 2: 128 pickup ticks on the evaluation clock of the one-warp Nelder-Mead instances: parity of the asynchronous forms, C3 / C4 lines
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "asynchronous or free_running" > $O/r2s3o_tests.log 2>&1; echo "tests rc=$?"; tail -3 $O/r2s3o_tests.log
timeout 600 python bench.py --workload C3 --steps 3 --warmup 3 --no-extra --no-cpu > $O/r2s3o_bench_c3.json 2> $O/r2s3o_bench_c3.err; echo "c3 rc=$?"
timeout 600 python bench.py --workload C4 --steps 3 --warmup 3 --no-extra --no-cpu > $O/r2s3o_bench_c4.json 2> $O/r2s3o_bench_c4.err; echo "c4 rc=$?"
python - <<'PY'
import json
for f in ("_c3", "_c4"):
    d = json.loads(open(f"gpurun_out/r2s3o_bench{f}.json").read().strip().splitlines()[-1])
    print(f, d["restarts_per_s"], d["stage_ms"], d.get("local_kernels_ms"), d["instances"], d["best_fval"], d.get("free_running"))
PY
