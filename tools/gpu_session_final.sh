#!/bin/bash
# round-1 final measurement session (1 GPU): tests, benches (C2 default + reference arm, C3/C4/C5), launch-floor
# microbenchmark, ncu launch list + full captures of the Sobol kernel at C2/C3/C4 and the restart kernel at C2
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python bench.py > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; tail -c 300 gpurun_out/bench_c2.err
timeout 300 python bench.py --impl reference --steps 3 > gpurun_out/bench_c2_ref.json 2>> gpurun_out/bench_c2.err
timeout 300 python bench.py --workload C3 --no-cpu --steps 3 > gpurun_out/bench_c3.json 2>&1
timeout 300 python bench.py --workload C4 --no-cpu --steps 3 > gpurun_out/bench_c4.json 2>&1
timeout 400 python bench.py --workload C5 --no-cpu --steps 2 --warmup 3 > gpurun_out/bench_c5.json 2> gpurun_out/bench_c5.err
./tools/micro/launch_floor > gpurun_out/launch_floor.txt 2>&1
timeout 300 python bench.py --no-cpu --steps 2 --warmup 3 > gpurun_out/plain_c2.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 120 -c 200 --csv --log-file gpurun_out/launches_c2.csv python bench.py --no-cpu --steps 2 --warmup 3 > gpurun_out/ncu_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"sobol_eval|nm_quad" -s 6 -c 3 -o gpurun_out/prof_c2 -f python bench.py --no-cpu --steps 2 --warmup 3 > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
timeout 300 python bench.py --workload C3 --no-local --no-cpu --steps 1 --warmup 3 > /dev/null 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"sobol_eval" -c 1 -o gpurun_out/prof_c3 -f python bench.py --workload C3 --no-local --no-cpu --steps 1 --warmup 3 > gpurun_out/ncu_full_c3.log 2>&1
timeout 300 python bench.py --workload C4 --no-local --no-cpu --steps 1 --warmup 3 > /dev/null 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"sobol_eval" -c 1 -o gpurun_out/prof_c4 -f python bench.py --workload C4 --no-local --no-cpu --steps 1 --warmup 3 > gpurun_out/ncu_full_c4.log 2>&1
nproc; nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv
