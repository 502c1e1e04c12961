"""experiment: time of the DFNLS stage on the C5 objective (full panel), a reduced job: 2000 Sobol points, 296 restarts"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tiktak_b200 import TikTakSearch, smm_config
K = int(os.environ.get("K", 295))
cfg = smm_config(qr_ndraw=2000, legitimate=1000, maxpoints=K, round_size=148)
with TikTakSearch(cfg) as s:
    s.solveAtSobolPoints(); s.chooseSobol()
    for rep in range(2):
        t0 = time.time(); s.LocalMinimizations("b"); torch.cuda.synchronize(); dt = time.time() - t0
        print(os.environ.get("TIKTAK_CUDA_LIB", "default").split("/")[-1], "K+1 =", K + 1, "local wall %.3f s" % dt, "stage_ms %.1f" % s.stage_ms("local"),
              "evals", int(s.evals.sum()), "rescues", int(s.nrescue.sum()) if hasattr(s, "nrescue") else None, flush=True)
