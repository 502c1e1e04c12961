#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
( time timeout 900 python bench.py > $O/r2h_bench.json 2> $O/r2h_bench.err ) 2> $O/r2h_time.txt; echo "bench rc=$?"
( time timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/r2h_ref.json 2> $O/r2h_ref.err ) 2>> $O/r2h_time.txt; echo "ref rc=$?"
cat $O/r2h_time.txt; tail -3 $O/r2h_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2h_bench.json'))
print('value %.3g'%d['value'],'frac %.3f'%d['roofline']['frac'],'restarts/s %.0f'%d['restarts_per_s'],d['stage_ms'],'e2e %.3g'%d['e2e']['value'], 'copy_tail', d['e2e']['copy_tail_s'], 'step_s', d['e2e']['step_s'])
for k,v in d['per_workload'].items(): print(k, 'evals/s %.3g'%v['evals_per_s'], 'frac %.3f'%v['roofline']['frac'], 'restarts/s %.0f'%v['restarts_per_s'], v['stage_ms'], 'cpu', v['cpu_baseline'] and ('%.3g'%v['cpu_baseline']['value'], '%.0f'%v['cpu_baseline']['restarts_per_s']))
print('parity', d['parity']['ok'], 'cpu', d['cpu_baseline']['value'], d['cpu_baseline']['restarts_per_s'], d['clocks'])
r=json.load(open('gpurun_out/r2h_ref.json')); print('ref', r['value'], r['steps'], r['ms_per_step'], r['config']['workload']==d['config']['workload'])
PY
