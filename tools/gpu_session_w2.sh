#!/bin/bash
# two-GPU check of the final kernels: sharded parity against the oracle, then the bench at N = 2
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 tests/mgpu_check.py 2>&1 | grep -E "world|MGPU|Error|error|ok|OK" | tail -6
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu > gpurun_out/scale_c2_g2.json 2> gpurun_out/scale_c2_g2.err
grep -E "Error|error" gpurun_out/scale_c2_g2.err | tail -2
python -c "
import json; d=json.loads(open('gpurun_out/scale_c2_g2.json').read().strip().splitlines()[-1]); print('N=2 value %.4g'%d['value'], 'restarts/s %.4g'%d['restarts_per_s'], d['stage_ms'], 'e2e %.4g'%d['e2e']['value'])"
