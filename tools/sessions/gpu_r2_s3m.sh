#!/bin/bash
# third session of round 2: find the free-running launches of nm_w4_kernel in the C3 command (by duration), then capture one in full
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
C3="python bench.py --workload C3 --steps 1 --warmup 3 --no-cpu --no-extra"
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -k regex:nm_w4_kernel -c 16 --csv --log-file $O/r2s3m_w4_list.csv $C3 > $O/r2s3m_ncu1.log 2>&1; echo "list rc=$?"
SKIP=$(python - <<'PY'
import csv
rows = [r for r in csv.reader(open("gpurun_out/r2s3m_w4_list.csv")) if len(r) > 5]
h = rows[0]; iid, im, iv = h.index("ID"), h.index("Metric Name"), h.index("Metric Value")
d = {}
for r in rows[1:]:
    if r[im] == "gpu__time_duration.sum": d.setdefault(r[iid], float(r[iv].replace(",", "")))
ids = sorted(d, key=int)
print("durations", [round(d[i]) for i in ids], file=__import__("sys").stderr)
big = max(d.values())
free = [n for n, i in enumerate(ids) if d[i] < 0.75 * big]
print(free[1] if len(free) > 1 else (free[0] if free else -1))
PY
)
echo "skip=$SKIP"
if [ "$SKIP" -ge 0 ]; then
timeout 900 ncu --set full --clock-control none --import-source on -k regex:nm_w4_kernel -s $SKIP -c 1 -o /tmp/r2s3m_c3_w4free $C3 > $O/r2s3m_ncu2.log 2>&1; echo "c3 w4 free rc=$?"
python tools/ncu_summary.py full /tmp/r2s3m_c3_w4free.ncu-rep $O/r02c_full_c3_nm_w4_free.csv "round 2 third session, final build: nm_w4_kernel as 4096 FREE-RUNNING one-warp instances (no virtual clock), C3 (10 001 restarts), one launch"
ncu -i /tmp/r2s3m_c3_w4free.ncu-rep --page raw --csv > $O/r2s3m_c3_w4free_raw.csv 2>/dev/null
sed -n 4,8p $O/r02c_full_c3_nm_w4_free.csv; grep -n "issue_active\|warps_active\|inst_executed.sum" $O/r02c_full_c3_nm_w4_free.csv
fi
