"""experiment: where the instances of an asynchronous run spend their time (TIKTAK_ASYNC_DEBUG=1)"""
import sys, os, ctypes as C
os.environ["TIKTAK_ASYNC_DEBUG"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tiktak_b200 import TikTakSearch, baseline_config
for w, P in (("C2", 256), ("C3", 888)):
    cfg = baseline_config(w, round_size=P)
    if w == "C3":
        cfg.qr_ndraw = 2_000_000
    with TikTakSearch(cfg) as s:
        s.solveAtSobolPoints(); s.chooseSobol()
        for rep in range(2):
            s.LocalMinimizationsAsync("a", instances=P); torch.cuda.synchronize()
        n = s.instances_used
        buf = (C.c_longlong * (5 * n))()
        s._L.tiktak_cuda_debug_async(C.c_void_p(s._h) if not isinstance(s._h, C.c_void_p) else s._h, buf, n)
        v = np.array(list(buf), dtype=np.float64).reshape(n, 5)
        ms = s.stage_ms("local")
        tot = v[:, :3].sum(1)
        print(w, "P", n, "local_ms %.3f" % ms, "cycles per instance: wait %.3g scan %.3g run %.3g (sum %.3g = %.2f ms at 1.9 GHz)" % (v[:, 0].mean(), v[:, 1].mean(), v[:, 2].mean(), tot.mean(), tot.mean() / 1.9e6),
              "restarts/instance %.1f" % v[:, 3].mean(), "run cycles per tick over the instances: min %.0f p10 %.0f median %.0f p90 %.0f max %.0f" % tuple(np.percentile(v[:, 2] / v[:, 4], [0, 10, 50, 90, 100])), "run cycles per restart %.3g" % (v[:, 2].sum() / v[:, 3].sum()),
              "wait per restart %.3g scan per restart %.3g" % (v[:, 0].sum() / v[:, 3].sum(), v[:, 1].sum() / v[:, 3].sum()), flush=True)
