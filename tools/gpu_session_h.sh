#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q -k "outside" 2>&1 | tail -4
run() { # lib shape workload group
  TIKTAK_CUDA_LIB=$1 TIKTAK_SOBOL_SHAPE=$2 TIKTAK_SOBOL_GROUP=$4 timeout 120 python bench.py --workload $3 --no-cpu --no-local --steps 8 2>/dev/null | python -c "
import json,sys
try:
    j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$3', '$(basename $1)', 'shape=$2 group=$4', '%.4g'%j['value'], 'frac %.3f'%j['roofline']['frac'], 'kernel_ms %.4f'%j['stage_ms']['pretest_eval_kernel'], 'pretest_ms %.4f'%j['stage_ms']['pretest'])
except Exception as e: print('$3 $1 $2 FAILED', e)"
}
D=$PWD/tiktak_b200/libtiktak_cuda.so
run $D auto C2 1
for v in mb7_g5 mb7_g10 mb8_g5; do for sh in 128x4 128x8; do run $PWD/tiktak_b200/libtiktak_cuda_$v.so $sh C2 1; done; done
run $D auto C2 1
./tools/micro/launch_floor | head -3
