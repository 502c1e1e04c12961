#!/bin/bash
# last measurement of round 1 (final kernels): default bench + reference arm, C3/C4 benches, ncu launch list and full
# captures of the Sobol kernel at C2 and C3
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python bench.py > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err
timeout 300 python bench.py --impl reference --steps 3 > gpurun_out/bench_c2_ref.json 2>> gpurun_out/bench_c2.err
timeout 300 python bench.py --workload C3 --no-cpu --steps 3 > gpurun_out/bench_c3.json 2>&1
timeout 300 python bench.py --workload C4 --no-cpu --steps 3 > gpurun_out/bench_c4.json 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 120 -c 200 --csv --log-file gpurun_out/launches_c2.csv python bench.py --no-cpu --steps 2 --warmup 3 > gpurun_out/ncu_launches.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"sobol_eval" -s 3 -c 1 -o gpurun_out/prof_c2_sobol -f python bench.py --no-cpu --no-local --steps 1 --warmup 3 > gpurun_out/ncu_c2s.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"sobol_eval" -c 1 -o gpurun_out/prof_c3 -f python bench.py --workload C3 --no-local --no-cpu --steps 1 --warmup 3 > gpurun_out/ncu_full_c3.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"sobol_eval" -c 1 -o gpurun_out/prof_c4 -f python bench.py --workload C4 --no-local --no-cpu --steps 1 --warmup 3 > gpurun_out/ncu_full_c4.log 2>&1
python - <<'PY'
import json
for n in ["bench_c2","bench_c2_ref","bench_c3","bench_c4"]:
    try:
        j=json.loads(open(f"gpurun_out/{n}.json").read().strip().splitlines()[-1])
        print(n, "value %.4g"%j["value"], "ms/step %.3f"%j["ms_per_step"], "restarts/s", j.get("restarts_per_s"), j.get("stage_ms"), "frac", (j.get("roofline") or {}).get("frac"), "e2e %.4g"%j["e2e"]["value"])
    except Exception as e: print(n, "ERR", e)
PY
