"""experiment: restart stage with asynchronous instances against the number of instances (TIKTAK_ASYNC_CB = pickup ticks from the environment)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tiktak_b200 import TikTakSearch, baseline_config
for w, Ps in (("C2", (64, 128, 256, 512)), ("C3", (296, 444, 592, 740, 887))):
    cfg = baseline_config(w, round_size=128)
    if w == "C3":
        cfg.qr_ndraw = 2_000_000
    with TikTakSearch(cfg) as s:
        s.solveAtSobolPoints(); s.chooseSobol()
        for P in Ps:
            ts = []
            for rep in range(3):
                torch.cuda.synchronize(); s.LocalMinimizationsAsync("a", instances=P); torch.cuda.synchronize()
                ts.append(s.stage_ms("local"))
            d = np.diff([0.0] + ts) if ts[1] > ts[0] * 1.5 else np.array(ts)
            print(w, "cb", os.environ.get("TIKTAK_ASYNC_CB", "32"), "P", s.instances_used, "local_ms", np.round(d, 3), "evals", int(s.evals.sum()), "best %.6g" % s.getBestLine()[0], flush=True)
