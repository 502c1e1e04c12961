#!/bin/bash
# last call of round 1: GPU tests + the bench lines kept under profiles/ (final code)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 200 python bench.py > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err
timeout 200 python bench.py --impl reference --steps 3 > gpurun_out/bench_c2_ref.json 2>> gpurun_out/bench_c2.err
timeout 200 python bench.py --workload C3 --no-cpu --steps 3 > gpurun_out/bench_c3.json 2>&1
timeout 200 python bench.py --workload C4 --no-cpu --steps 2 > gpurun_out/bench_c4.json 2>&1
python - <<'PY'
import json
for n in ["bench_c2","bench_c2_ref","bench_c3","bench_c4"]:
    try:
        j=json.loads(open(f"gpurun_out/{n}.json").read().strip().splitlines()[-1])
        print(n, "value %.4g"%j["value"], "ms/step %.3f"%j["ms_per_step"], "restarts/s", j.get("restarts_per_s"), j.get("stage_ms"), "frac", (j.get("roofline") or {}).get("frac"), "e2e %.4g"%j["e2e"]["value"])
    except Exception as e: print(n, "ERR", e)
PY
