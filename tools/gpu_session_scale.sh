#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29521 tools/mgpu_check.py 2>&1 | grep -E "world|MGPU" 
for n in 1 2 4 8; do
  if [ $n -eq 1 ]; then python bench.py --gpus 1 --steps 3 --warmup 3 --no-cpu > gpurun_out/scale_c2_g$n.json 2> gpurun_out/scale_c2_g$n.err
  else python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29530+n)) bench.py --gpus $n --steps 3 --warmup 3 > gpurun_out/scale_c2_g$n.json 2> gpurun_out/scale_c2_g$n.err; fi
  tail -c 250 gpurun_out/scale_c2_g$n.err | tail -2
done
for n in 1 8; do
  if [ $n -eq 1 ]; then python bench.py --gpus 1 --workload C3 --steps 2 --warmup 3 --no-cpu --round-size 1024 > gpurun_out/scale_c3_g$n.json 2> gpurun_out/scale_c3_g$n.err
  else python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29550+n)) bench.py --gpus $n --workload C3 --steps 2 --warmup 3 --round-size 1024 > gpurun_out/scale_c3_g$n.json 2> gpurun_out/scale_c3_g$n.err; fi
done
python - <<'PY'
import json
for w in ("c2","c3"):
    for n in (1,2,4,8):
        try:
            d=json.loads(open(f"gpurun_out/scale_{w}_g{n}.json").read().strip().splitlines()[-1])
            print(w, n, "value %.4g"%d["value"], "restarts/s", d["restarts_per_s"], "ms/step", d["ms_per_step"], d["stage_ms"])
        except Exception as e: print(w, n, "ERR", e)
PY
