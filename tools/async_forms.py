"""experiment: the restart stage by asynchronous instances, four-warp blocks (TIKTAK_ASYNC_FORM=quad) against one-warp instances (w4),
over the number of instances; rounds beside it where asked"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tiktak_b200 import TikTakSearch, baseline_config

def run(w, form, Ps, mode=None, wpb=None):
    os.environ["TIKTAK_ASYNC_FORM"] = form
    for k, v in (("TIKTAK_ASYNC_MODE", mode), ("TIKTAK_ASYNC_WPB", wpb)):
        if v: os.environ[k] = str(v)
        else: os.environ.pop(k, None)
    cfg = baseline_config(w, round_size=128)
    if w in ("C3", "C4"):
        cfg.qr_ndraw = 2_000_000
    with TikTakSearch(cfg) as s:
        s.solveAtSobolPoints(); s.chooseSobol()
        for P in Ps:
            ts = []
            for rep in range(3):
                torch.cuda.synchronize(); s.LocalMinimizationsAsync("a", instances=P); torch.cuda.synchronize()
                ts.append(s.stage_ms("local"))
            print(w, form, "mode", mode, "wpb", wpb, "P", s.instances_used, "local_ms", np.round(ts, 3), "evals", int(s.evals.sum()), "best %.6g" % s.getBestLine()[0], flush=True)

which = sys.argv[1:] or ["C2", "C3", "C4"]
if "C2" in which:
    run("C2", "quad", (256, 512))
    run("C2", "w4", (256, 512, 1001))
    run("C2", "w4", (512, 1001), mode=324)
if "C3" in which:
    run("C3", "quad", (887,))
    run("C3", "w4", (592, 1184, 1776, 2368, 2959, 4000))
    run("C3", "w4", (1184, 2368, 4000), mode=324)
    run("C3", "w4", (1184, 2368), wpb=1)
if "C4" in which:
    run("C4", "w4", (1184, 2368, 3552, 5000))
