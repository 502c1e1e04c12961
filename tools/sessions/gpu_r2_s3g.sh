#!/bin/bash
# third session of round 2, after the fetch was rewritten (rows packed in place, no host transposition): all GPU tests, bench.py as the
# driver runs it with the reference arm, the C3 / C4 lines, smoke
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > $O/r2s3g_tests.log 2>&1; echo "tests rc=$?"; tail -3 $O/r2s3g_tests.log
( time timeout 900 python bench.py --steps 5 --warmup 3 > $O/r2s3g_bench.json 2> $O/r2s3g_bench.err ) 2> $O/r2s3g_time.txt; echo "bench rc=$?"; tail -3 $O/r2s3g_time.txt; tail -2 $O/r2s3g_bench.err
( time timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/r2s3g_bench_ref.json 2> $O/r2s3g_bench_ref.err ) 2> $O/r2s3g_time_ref.txt; echo "ref rc=$?"; tail -3 $O/r2s3g_time_ref.txt
python -c "import __graft_entry__ as g; g.smoke()" > $O/r2s3g_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $O/r2s3g_smoke.log
timeout 600 python bench.py --workload C3 --steps 3 --warmup 3 --no-extra > $O/r2s3g_bench_c3.json 2> $O/r2s3g_bench_c3.err; echo "c3 rc=$?"
timeout 600 python bench.py --workload C4 --steps 3 --warmup 3 --no-extra > $O/r2s3g_bench_c4.json 2> $O/r2s3g_bench_c4.err; echo "c4 rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2s3g_bench.json").read().strip().splitlines()[-1])
print("value %.3g frac %.3f e2e %.3g" % (d["value"], d["roofline"]["frac"], d["e2e"]["value"]), d.get("parity", {}).get("ok"), d.get("restarts_per_s"), d.get("stage_ms"), d.get("rounds"))
for k, v in (d.get("per_workload") or {}).items():
    print(k, {kk: v.get(kk) for kk in ("restarts_per_s", "stage_ms", "local_kernels_ms", "instances", "best_fval")})
for f in ("c3", "c4"):
    d = json.loads(open(f"gpurun_out/r2s3g_bench_{f}.json").read().strip().splitlines()[-1])
    print(f, d["restarts_per_s"], d["stage_ms"], d["instances"], d["best_fval"], d.get("rounds"), d.get("four_warp_blocks"), d.get("free_running"), d.get("roofline_local", {}).get("frac"))
PY
