"""experiment: the restart stage at BASELINE sizes, rounds (LocalMinimizations) against asynchronous instances (LocalMinimizationsAsync)"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tiktak_b200 import TikTakSearch, baseline_config
for w, B, Ps in (("C2", 128, (128, 256, 888)), ("C3", 1024, (888,))):
    cfg = baseline_config(w, round_size=B)
    if w == "C3":
        cfg.qr_ndraw = 2_000_000  # the restart stage does not depend on how the starts were found: a smaller pre-test
    with TikTakSearch(cfg) as s:
        s.solveAtSobolPoints(); s.chooseSobol()
        for rep in range(2):
            s.LocalMinimizations("a"); torch.cuda.synchronize()
            print(w, "rounds of", B, "local_ms %.3f" % s.stage_ms("local"), "launches", s.local_launches, "evals", int(s.evals.sum()), "best %.6g" % s.getBestLine()[0], flush=True)
        for P in Ps:
            for rep in range(2):
                t0 = time.time(); s.LocalMinimizationsAsync("a", instances=P); torch.cuda.synchronize(); dt = time.time() - t0
                print(w, "async P", P, "used", s.instances_used, "local_ms %.3f" % s.stage_ms("local"), "evals", int(s.evals.sum()), "best %.6g" % s.getBestLine()[0], flush=True)
