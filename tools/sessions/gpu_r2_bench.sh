#!/bin/bash
# the GPU test suite, then bench.py as the driver runs it (N = 1) and the reference arm
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > $O/r2b_tests.log 2>&1; echo "tests rc=$?"; tail -3 $O/r2b_tests.log
( time timeout 900 python bench.py --steps 5 --warmup 3 > $O/r2b_bench.json 2> $O/r2b_bench.err ) 2> $O/r2b_time.txt; echo "bench rc=$?"; tail -3 $O/r2b_time.txt; tail -3 $O/r2b_bench.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2b_bench.json') if l.startswith('{')][-1])
print('value %.3g'%d['value'], 'frac %.3f'%d['roofline']['frac'], 'restarts/s %.0f'%d['restarts_per_s'], d['local_mode'], d['instances'], d['stage_ms'], 'rounds', d['rounds'], 'e2e %.3g'%d['e2e']['value'], 'parity', d['parity']['ok'], 'launches', d['gpu_launches'])
for k,v in (d['per_workload'] or {}).items():
    print(k, 'frac %.3f'%v['roofline']['frac'], 'restarts/s %.0f'%v['restarts_per_s'], v['local_mode'], v['instances'], {a: round(b,3) for a,b in v['stage_ms'].items()}, 'rounds', v.get('rounds'), 'best', v['best_fval'])
print(json.dumps(d['parity']['cases']))
PY
