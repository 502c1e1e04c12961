#!/bin/bash
# round-2 session A: parity of the throughput Nelder-Mead form + first timings of both forms
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > $O/r2a_pytest_tp.log 2>&1; echo "pytest tp rc=$?" > $O/r2a_status.txt
for w in C2 C3 C4; do
  timeout 300 python bench.py --workload $w --steps 3 --warmup 3 --no-cpu > $O/r2a_bench_${w}_tp.json 2> $O/r2a_bench_${w}_tp.err; echo "bench $w tp rc=$?" >> $O/r2a_status.txt
  TIKTAK_NM_FORM=quad timeout 300 python bench.py --workload $w --steps 3 --warmup 3 --no-cpu > $O/r2a_bench_${w}_quad.json 2> $O/r2a_bench_${w}_quad.err; echo "bench $w quad rc=$?" >> $O/r2a_status.txt
done
tail -5 $O/r2a_pytest_tp.log; cat $O/r2a_status.txt
