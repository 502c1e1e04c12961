#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
W=${1:-C2}; export TIKTAK_NM_LP=${2:-32}
timeout 300 python bench.py --workload $W --steps 1 --warmup 3 --no-cpu > $O/r2c_plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:nm_w4_kernel -s 12 -c 1 -o $O/r2c_${W}_w4 python bench.py --workload $W --steps 1 --warmup 3 --no-cpu > $O/r2c_ncu.log 2>&1
echo "rc=$?"; tail -2 $O/r2c_ncu.log | cut -c1-300
