#!/bin/bash
# experiment: Sobol kernel launch shapes (threads per block x points per thread) at C2 / C4
cd "$(dirname "$0")/.."
for w in C2 C4; do
for shape in default 64x4 64x8 64x16 128x2 128x4 128x8 128x16 256x4 256x8; do
  if [ $shape = default ]; then unset TIKTAK_SOBOL_SHAPE; else export TIKTAK_SOBOL_SHAPE=$shape; fi
  timeout 120 python bench.py --workload $w --no-local --no-cpu --steps 6 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$w $shape', 'kernel_ms %.4f'%d['stage_ms']['pretest_eval_kernel'], 'stage_ms %.4f'%d['stage_ms']['pretest'], 'frac %.3f'%d['roofline']['frac'])"
done; done
