#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --no-cpu > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; tail -c 300 gpurun_out/bench_c2.err
python bench.py --workload C5 --steps 2 --warmup 3 --no-cpu > gpurun_out/bench_c5.json 2> gpurun_out/bench_c5.err; tail -c 300 gpurun_out/bench_c5.err
python - <<'PY'
import json
for f in ["gpurun_out/bench_c2.json","gpurun_out/bench_c5.json"]:
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print(f, "value %.4g"%d["value"], "restarts/s", d["restarts_per_s"], d["stage_ms"], "frac", d["roofline"]["frac"], "e2e %.4g"%d["e2e"]["value"], d["e2e"]["step_s"])
PY
