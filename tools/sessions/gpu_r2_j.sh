#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
for v in base mb7 mb8; do
  lib=$PWD/tiktak_b200/libtiktak_cuda.so; [ $v != base ] && lib=$PWD/tiktak_b200/libtiktak_cuda_$v.so
  for w in C2 C3 C4; do
    TIKTAK_CUDA_LIB=$lib timeout 300 python bench.py --workload $w --no-extra --no-cpu --no-local --steps 20 --warmup 5 > $O/r2j_${v}_$w.json 2> $O/r2j_${v}_$w.err
    python - $O/r2j_${v}_$w.json $v $w <<'PY'
import json,sys
d=json.load(open(sys.argv[1])); s=sorted(d['pretest_ms_steps'])
print(sys.argv[2], sys.argv[3], 'pretest ms mean %.5f median %.5f min %.5f'%(d['stage_ms']['pretest'], s[len(s)//2], s[0]), 'frac %.3f'%d['roofline']['frac'])
PY
  done
done
