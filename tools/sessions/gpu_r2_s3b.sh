#!/bin/bash
# third session of round 2: bench as the driver runs it, reference arm, launch list, full capture of the one-warp asynchronous kernel (C3)
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
( time timeout 900 python bench.py --steps 5 --warmup 3 > $O/r2s3_bench.json 2> $O/r2s3_bench.err ) 2> $O/r2s3_time.txt; echo "bench rc=$?"; tail -3 $O/r2s3_time.txt; tail -2 $O/r2s3_bench.err
( time timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/r2s3_bench_ref.json 2> $O/r2s3_bench_ref.err ) 2> $O/r2s3_time_ref.txt; echo "ref rc=$?"; tail -3 $O/r2s3_time_ref.txt
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2s3_bench.json").read().strip().splitlines()[-1])
print("value %.3g frac %.3f e2e %.3g" % (d["value"], d["roofline"]["frac"], d["e2e"]["value"]), d.get("parity", {}).get("ok"), d.get("restarts_per_s"))
for k, v in (d.get("per_workload") or {}).items():
    print(k, {kk: v.get(kk) for kk in ("restarts_per_s", "instances", "instance_form", "best_fval", "rounds", "four_warp_blocks", "free_running")})
PY
C3="python bench.py --workload C3 --steps 1 --warmup 3 --no-cpu --no-extra"
timeout 600 python bench.py --workload C3 --steps 3 --warmup 3 --no-extra > $O/r2s3_bench_c3.json 2> $O/r2s3_bench_c3.err; echo "c3 rc=$?"
timeout 600 python bench.py --workload C4 --steps 3 --warmup 3 --no-extra > $O/r2s3_bench_c4.json 2> $O/r2s3_bench_c4.err; echo "c4 rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:nm_w4_kernel -s 3 -c 1 -o $O/r2s3_c3_w4async $C3 > $O/r2s3_ncu2.log 2>&1; echo "c3 w4 async rc=$?"
ls -la $O/r2s3_*
