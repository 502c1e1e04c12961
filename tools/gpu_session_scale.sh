#!/bin/bash
# scaling session on an 8-GPU box: sharded parity check at world 8, then bench.py at N = 1, 2, 4, 8 (C2) and C3 at 8
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 tests/mgpu_check.py 2>&1 | grep -E "world|MGPU|Error|error" | tail -6
for n in 1 2 4 8; do
  if [ $n -eq 1 ]; then timeout 300 python bench.py --gpus 1 --steps 3 --warmup 3 --no-cpu > gpurun_out/scale_c2_g$n.json 2> gpurun_out/scale_c2_g$n.err
  else timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29530+n)) bench.py --gpus $n --steps 3 --warmup 3 > gpurun_out/scale_c2_g$n.json 2> gpurun_out/scale_c2_g$n.err; fi
  grep -E "Error|error" gpurun_out/scale_c2_g$n.err | tail -2
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29558 bench.py --gpus 8 --workload C3 --steps 2 --warmup 3 > gpurun_out/scale_c3_g8.json 2> gpurun_out/scale_c3_g8.err
python - <<'PY'
import json
for w, ns in (("c2", (1, 2, 4, 8)), ("c3", (8,))):
    for n in ns:
        try:
            d = json.loads(open(f"gpurun_out/scale_{w}_g{n}.json").read().strip().splitlines()[-1])
            print(w, n, "value %.4g" % d["value"], "restarts/s %.4g" % (d["restarts_per_s"] or 0), "ms/step %.3f" % d["ms_per_step"], d["stage_ms"], "e2e %.4g" % d["e2e"]["value"])
        except Exception as e:
            print(w, n, "ERR", e)
PY
