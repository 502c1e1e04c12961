#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
./tools/micro/lat > $O/r2b_lat.txt 2>&1
timeout 300 python bench.py --workload C3 --steps 1 --warmup 3 --no-cpu > $O/r2b_plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:nm_tp_kernel -s 10 -c 1 -o $O/r2b_c3_tp python bench.py --workload C3 --steps 1 --warmup 3 --no-cpu > $O/r2b_ncu.log 2>&1
echo "rc=$?"; cat $O/r2b_lat.txt; tail -3 $O/r2b_ncu.log
