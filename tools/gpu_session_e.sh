#!/bin/bash
# experiment: register budget (MB) x group width (G) x points per thread (P) of the grouped Sobol kernel
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { # lib shape workload
  TIKTAK_CUDA_LIB=$1 TIKTAK_SOBOL_SHAPE=$2 timeout 120 python bench.py --workload $3 --no-cpu --no-local --steps 5 2>/dev/null | python -c "
import json,sys
try:
    j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$3', '$(basename $1)', '$2', '%.4g'%j['value'], 'frac %.3f'%j['roofline']['frac'], 'kernel_ms %.4f'%j['stage_ms']['pretest_eval_kernel'])
except Exception as e: print('$3 $1 $2 FAILED', e)"
}
TIKTAK_CUDA_LIB=$PWD/tiktak_b200/libtiktak_cuda_mb4_g10.so timeout 300 python -m pytest tests -m gpu -x -q -k "pretest or full_size or legitimate or choose or smoke" 2>&1 | tail -3
for v in mb3_g10 mb4_g10 mb4_g5 mb5_g10 mb6_g10 mb6_g5; do
  L=$PWD/tiktak_b200/libtiktak_cuda_$v.so
  for sh in 128x2 128x4 128x8 128x16; do run $L $sh C2; done
  run $L auto C3
done
run $PWD/tiktak_b200/libtiktak_cuda.so 128x2 C2
