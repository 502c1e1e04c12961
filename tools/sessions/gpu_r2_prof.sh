#!/bin/bash
# round-2 profiles: launch list of the default bench command, full captures of the restart kernels (C2 quad, C4 one-warp, C5 DFNLS)
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu --no-extra"
timeout 300 $CMD > $O/r2p_plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 400 --csv --log-file $O/r2p_launches_c2.csv $CMD > $O/r2p_ncu1.log 2>&1
echo "launch list rc=$?"
C4="python bench.py --workload C4 --steps 1 --warmup 3 --no-cpu --no-extra"
timeout 300 $C4 > $O/r2p_plain4.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:nm_w4_kernel -s 20 -c 1 -o $O/r2p_c4_w4 $C4 > $O/r2p_ncu2.log 2>&1
echo "c4 w4 rc=$?"
C5="python bench.py --workload C5 --steps 1 --warmup 3 --no-cpu --no-extra"
timeout 600 $C5 > $O/r2p_plain5.log 2>&1 &&
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:dfnls_kernel -s 4 -c 1 -o $O/r2p_c5_dfnls $C5 > $O/r2p_ncu3.log 2>&1
echo "c5 dfnls rc=$?"
ls -la $O/r2p_*
