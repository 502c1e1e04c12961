#!/bin/bash
cd "$(dirname "$0")/.."
run() { TIKTAK_CUDA_LIB=$PWD/tiktak_b200/$1 timeout 200 python bench.py --workload $2 --no-cpu --steps 5 2>/dev/null | python -c "
import json,sys
j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1 $2', 'local_ms %.4f'%j['stage_ms']['local'], 'restarts/s %.4g'%j['restarts_per_s'], 'pretest %.4f'%j['stage_ms']['pretest'])"; }
for i in 1 2; do run libtiktak_cuda.so C2; run libtiktak_cuda_plain.so C2; done
run libtiktak_cuda.so C3; run libtiktak_cuda_plain.so C3
