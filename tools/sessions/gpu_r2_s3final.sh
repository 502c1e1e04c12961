#!/bin/bash
# third session of round 2, final numbers: GPU tests, bench.py as the driver runs it (N = 1) with the reference arm, C3 / C4 lines, the ncu
# launch list of the default command and full captures of the one-warp asynchronous kernel at C3 and C4
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > $O/r2s3f_tests.log 2>&1; echo "tests rc=$?"; tail -3 $O/r2s3f_tests.log
( time timeout 900 python bench.py --steps 5 --warmup 3 > $O/r2s3f_bench.json 2> $O/r2s3f_bench.err ) 2> $O/r2s3f_time.txt; echo "bench rc=$?"; tail -3 $O/r2s3f_time.txt; tail -2 $O/r2s3f_bench.err
( time timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/r2s3f_bench_ref.json 2> $O/r2s3f_bench_ref.err ) 2> $O/r2s3f_time_ref.txt; echo "ref rc=$?"; tail -3 $O/r2s3f_time_ref.txt
python -c "import __graft_entry__ as g; g.smoke()" > $O/r2s3f_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $O/r2s3f_smoke.log
timeout 600 python bench.py --workload C3 --steps 3 --warmup 3 --no-extra > $O/r2s3f_bench_c3.json 2> $O/r2s3f_bench_c3.err; echo "c3 rc=$?"
timeout 600 python bench.py --workload C4 --steps 3 --warmup 3 --no-extra > $O/r2s3f_bench_c4.json 2> $O/r2s3f_bench_c4.err; echo "c4 rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2s3f_bench.json").read().strip().splitlines()[-1])
print("value %.3g frac %.3f e2e %.3g" % (d["value"], d["roofline"]["frac"], d["e2e"]["value"]), d.get("parity", {}).get("ok"), d.get("restarts_per_s"), d.get("stage_ms"), d.get("local_kernels_ms"))
for k, v in (d.get("per_workload") or {}).items():
    print(k, {kk: v.get(kk) for kk in ("restarts_per_s", "stage_ms", "local_kernels_ms", "instances", "best_fval", "rounds", "four_warp_blocks", "free_running")})
for f in ("c3", "c4"):
    d = json.loads(open(f"gpurun_out/r2s3f_bench_{f}.json").read().strip().splitlines()[-1])
    print(f, d["restarts_per_s"], d["stage_ms"], d.get("local_kernels_ms"), d["instances"], d["best_fval"], d.get("rounds"), d.get("four_warp_blocks"), d.get("roofline_local", {}).get("frac"), d.get("cpu_baseline"))
PY
CMD="python bench.py --steps 2 --warmup 3 --no-cpu --no-extra"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 400 --csv --log-file $O/r2s3f_launches_c2.csv $CMD > $O/r2s3f_ncu1.log 2>&1; echo "launch list rc=$?"
C3="python bench.py --workload C3 --steps 1 --warmup 3 --no-cpu --no-extra"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:nm_w4_kernel -s 3 -c 1 -o $O/r2s3f_c3_w4async $C3 > $O/r2s3f_ncu2.log 2>&1; echo "c3 w4 async rc=$?"
C4="python bench.py --workload C4 --steps 1 --warmup 3 --no-cpu --no-extra"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:nm_w4_kernel -s 3 -c 1 -o $O/r2s3f_c4_w4async $C4 > $O/r2s3f_ncu3.log 2>&1; echo "c4 w4 async rc=$?"
ls -la $O/r2s3f_*
