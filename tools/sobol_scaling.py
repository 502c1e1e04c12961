"""experiment: Sobol evaluation kernel time vs number of points (intercept = fixed cost, slope = per-point cost)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tiktak_b200 import TikTakSearch, baseline_config
flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device="cuda")
for w in ("C2", "C3"):
    for n in (65535, 250000, 500000, 1000000, 2000000, 4000000, 8000000, 16000000):
        cfg = baseline_config(w); cfg.qr_ndraw = n; cfg.legitimate = n + 10000; cfg.maxpoints = 10
        with TikTakSearch(cfg) as s:
            ts = []
            for it in range(8):
                flush.zero_(); torch.cuda.synchronize()
                s.solveAtSobolPoints(wait=False); torch.cuda.synchronize()
                ts.append(s.stage_ms("pretest_eval"))
            print(w, n, "kernel_us min %.2f med %.2f" % (1e3 * min(ts[2:]), 1e3 * float(np.median(ts[2:]))), flush=True)
