#!/bin/bash
# asynchronous instances for DFNLS: parity tests, then the C5 restart stage at BASELINE size in both forms
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "asynchronous or dfnls" > $O/r2async2_tests.log 2>&1; echo "tests rc=$?"; tail -8 $O/r2async2_tests.log
timeout 600 python - <<'PY'
import sys; sys.path.insert(0, '.')
import torch
from tiktak_b200 import TikTakSearch, smm_config
cfg = smm_config(qr_ndraw=3000, legitimate=1500, maxpoints=1000, round_size=128)
with TikTakSearch(cfg) as s:
    s.solveAtSobolPoints(); s.chooseSobol()
    s.LocalMinimizations("b"); torch.cuda.synchronize()
    print("rounds of 128: local_ms %.1f" % s.stage_ms("local"), "evals", int(s.evals.sum()), "best %.6g" % s.getBestLine()[0], flush=True)
    for P in (128, 147):
        s.LocalMinimizationsAsync("b", instances=P); torch.cuda.synchronize()
        print("async P", s.instances_used, "local_ms %.1f" % s.stage_ms("local"), "evals", int(s.evals.sum()), "best %.6g" % s.getBestLine()[0], "rescues", int(s.nrescue.sum()), flush=True)
PY
