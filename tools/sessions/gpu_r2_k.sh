#!/bin/bash
# N-GPU bench exactly as the driver launches it
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
N=${1:-2}
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 5 --warmup 3 > $O/r2k_bench_g$N.json 2> $O/r2k_bench_g$N.err ) 2> $O/r2k_time_g$N.txt; echo "rc=$?"
tail -3 $O/r2k_time_g$N.txt; tail -2 $O/r2k_bench_g$N.err
python - $O/r2k_bench_g$N.json <<'PY'
import json,sys
d=json.loads([l for l in open(sys.argv[1]) if l.startswith('{')][-1])
print('n_gpus',d['n_gpus'],'value %.3g'%d['value'],'restarts/s %.0f'%d['restarts_per_s'],d['stage_ms'],'e2e %.3g'%d['e2e']['value'],'parity',d['parity']['ok'])
s=d['strong_c3']; print('strong_c3', s['stage_ms'], 'step ms %.2f'%s['ms_per_step'], 'restarts/s %.0f'%s['restarts_per_s'])
PY
