#!/bin/bash
# third session of round 2, last captures: launch list of the default command on the final build; full capture of a free-running one-warp launch (C3)
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu --no-extra"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 400 --csv --log-file $O/r2s3l_launches_c2.csv $CMD > $O/r2s3l_ncu1.log 2>&1; echo "launch list rc=$?"
C3="python bench.py --workload C3 --steps 1 --warmup 3 --no-cpu --no-extra"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:nm_w4_kernel -s 5 -c 1 -o /tmp/r2s3l_c3_w4free $C3 > $O/r2s3l_ncu2.log 2>&1; echo "c3 w4 free rc=$?"
python tools/ncu_summary.py full /tmp/r2s3l_c3_w4free.ncu-rep $O/r02c_full_c3_nm_w4_free.csv "round 2 third session: nm_w4_kernel as 4096 FREE-RUNNING one-warp instances (no virtual clock), C3 (10 001 restarts), one launch (the 6th nm_w4_kernel launch of: $C3)"
ncu -i /tmp/r2s3l_c3_w4free.ncu-rep --page raw --csv > $O/r2s3l_c3_w4free_raw.csv 2>/dev/null
cat $O/r02c_full_c3_nm_w4_free.csv
