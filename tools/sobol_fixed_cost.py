"""experiment: what the fixed cost of the Sobol evaluation kernel is made of -- kernel time (library events) for a
few sizes with the L2 flushed before every launch (cold code, constants) and without (warm)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tiktak_b200 import TikTakSearch, baseline_config
flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device="cuda")
for w in sys.argv[1:] or ("C2",):
    for n in [int(v) for v in os.environ.get("NS", "1000,16384,65535,250000,1000000,4000000").split(",")]:
        cfg = baseline_config(w); cfg.qr_ndraw = n; cfg.legitimate = n + 10000; cfg.maxpoints = 10
        with TikTakSearch(cfg) as s:
            for cold in (True, False):
                ts = []
                for it in range(10):
                    if cold:
                        flush.zero_()
                    torch.cuda.synchronize()
                    s.solveAtSobolPoints(wait=False); torch.cuda.synchronize()
                    ts.append(s.stage_ms("pretest_eval"))
                print(w, n, "cold" if cold else "warm", "kernel_us min %.2f med %.2f" % (1e3 * min(ts[2:]), 1e3 * float(np.median(ts[2:]))), flush=True)
