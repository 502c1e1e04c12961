#!/bin/bash
# second session of round 2, final numbers: GPU tests, bench.py as the driver runs it (N = 1) with the reference arm, the ncu launch list
# of the default command and full captures of the asynchronous Nelder-Mead kernel (C2) and the Sobol kernel inside the stage graph
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > $O/r2f_tests.log 2>&1; echo "tests rc=$?"; tail -3 $O/r2f_tests.log
( time timeout 900 python bench.py --steps 5 --warmup 3 > $O/r2f_bench.json 2> $O/r2f_bench.err ) 2> $O/r2f_time.txt; echo "bench rc=$?"; tail -3 $O/r2f_time.txt; tail -2 $O/r2f_bench.err
( time timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/r2f_bench_ref.json 2> $O/r2f_bench_ref.err ) 2> $O/r2f_time_ref.txt; echo "ref rc=$?"; tail -3 $O/r2f_time_ref.txt
python -c "import __graft_entry__ as g; g.smoke()" > $O/r2f_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $O/r2f_smoke.log
CMD="python bench.py --steps 2 --warmup 3 --no-cpu --no-extra"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 400 --csv --log-file $O/r2f_launches_c2.csv $CMD > $O/r2f_ncu1.log 2>&1; echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:nm_quad_async_kernel -s 3 -c 1 -o $O/r2f_c2_async $CMD > $O/r2f_ncu2.log 2>&1; echo "c2 async rc=$?"
C4="python bench.py --workload C4 --steps 1 --warmup 3 --no-cpu --no-extra"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:nm_w4_kernel -s 20 -c 1 -o $O/r2f_c4_w4 $C4 > $O/r2f_ncu3.log 2>&1; echo "c4 w4 rc=$?"
ls -la $O/r2f_*
