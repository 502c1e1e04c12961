#!/bin/bash
# measurement session: tests, benches (C2 default, C3/C4/C5), ncu launch list + full captures
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; tail -c 400 gpurun_out/bench_c2.err
python bench.py --impl reference --steps 3 > gpurun_out/bench_c2_ref.json 2>> gpurun_out/bench_c2.err
python bench.py --workload C5 --steps 2 --warmup 3 --cpu-seconds 20 > gpurun_out/bench_c5.json 2> gpurun_out/bench_c5.err; tail -c 600 gpurun_out/bench_c5.err
python bench.py --workload C4 --no-cpu --steps 2 > gpurun_out/bench_c4.json 2>&1
python bench.py --workload C3 --no-local --no-cpu --steps 2 > gpurun_out/bench_c3.json 2>&1
python bench.py --no-cpu --steps 2 --warmup 3 > gpurun_out/plain_c2.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 250 -c 300 --csv --log-file gpurun_out/launches_c2.csv python bench.py --no-cpu --steps 2 --warmup 3 > gpurun_out/ncu_launches.log 2>&1
python bench.py --no-cpu --steps 2 --warmup 3 > gpurun_out/plain_c2b.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"sobol_eval|nm_lane" -s 6 -c 4 -o gpurun_out/prof_c2 python bench.py --no-cpu --steps 2 --warmup 3 > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
nproc; nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv
