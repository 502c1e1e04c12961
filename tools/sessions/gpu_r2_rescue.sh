#!/bin/bash
# RESCUE_H on the device, DFNLS up to nx = 100: the GPU test suite, then C5 at BASELINE size (time of the DFNLS stage, RESCUE_H calls)
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > $O/r2resc_tests.log 2>&1; echo "tests rc=$?"; tail -5 $O/r2resc_tests.log
timeout 600 python bench.py --workload C5 --steps 1 --warmup 1 --no-cpu --no-extra > $O/r2resc_c5.json 2> $O/r2resc_c5.err; echo "c5 rc=$?"
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2resc_c5.json') if l.startswith('{')][-1])
print('C5 stage_ms', d['stage_ms'], 'restarts/s', d.get('restarts_per_s'))
PY
