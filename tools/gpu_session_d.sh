#!/bin/bash
# short session: GPU tests, then the Sobol stage at C2 / C3 / C4 (kernel shapes at C2)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for w in C2 C3 C4; do
  timeout 200 python bench.py --workload $w --no-cpu --no-local --steps 5 > gpurun_out/d_$w.json 2> gpurun_out/d_$w.err
  python - <<PY
import json
try:
    j=json.loads(open("gpurun_out/d_$w.json").read().strip().splitlines()[-1])
    print("$w", "%.4g evals/s" % j["value"], "frac %.3f" % j["roofline"]["frac"], j["stage_ms"], "e2e %.3g" % j["e2e"]["value"])
except Exception as e:
    print("$w failed", e); print(open("gpurun_out/d_$w.err").read()[-800:])
PY
done
for sh in 128x4 128x16; do
  TIKTAK_SOBOL_SHAPE=$sh timeout 200 python bench.py --workload C2 --no-cpu --no-local --steps 5 2>/dev/null | python -c "
import json,sys
j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$sh', '%.4g'%j['value'], j['roofline']['frac'], j['stage_ms'])"
done
