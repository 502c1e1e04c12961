#!/bin/bash
# the one-wave pre-test stage as a CUDA graph: GPU tests, then C2 / C3 / C4 with and without the graph
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > $O/r2graph_tests.log 2>&1; echo "tests rc=$?"; tail -3 $O/r2graph_tests.log
run() { # workload, label
  timeout 120 python bench.py --workload $1 --no-local --no-cpu --no-extra --steps 8 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1 $2', 'kernel_ms %.4f'%d['stage_ms']['pretest_eval_kernel'], 'stage_ms %.4f'%d['stage_ms']['pretest'], 'frac %.3f'%d['roofline']['frac'], 'e2e %.3g'%d['e2e']['value'])"
}
for rep in 1 2 3; do run C2 graph; TIKTAK_NO_GRAPH=1 run C2 stream; done
for w in C3 C4; do run $w graph; TIKTAK_NO_GRAPH=1 run $w stream; done
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu > $O/r2graph_bench.json 2> $O/r2graph_bench.err; echo "bench rc=$?"; tail -c 600 $O/r2graph_bench.json
