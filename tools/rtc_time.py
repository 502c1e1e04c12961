"""experiment: Sobol stage for (objective, nx) pairs without a compiled-in instantiation: the run-time compiled fixed-nx kernel
(default) against the runtime-nx kernel (TIKTAK_NO_RTC=1, set in the environment of a second run)"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tiktak_b200 import TikTakSearch, make_config, OBJ_GRIEWANK, OBJ_RASTRIGIN, OBJ_ROSENBROCK, lib
for name, obj, nx, lo, hi, fpc in (("griewank", OBJ_GRIEWANK, 7, -400.0, 600.0, 54), ("rastrigin", OBJ_RASTRIGIN, 16, -4.12, 5.12, 55), ("rosenbrock", OBJ_ROSENBROCK, 30, -5.0, 10.0, 18)):
    cfg = make_config(nx, 1, 4_000_000, 10, lo, hi, init=[0.5 * (lo + hi)] * nx, objective_id=obj)
    with TikTakSearch(cfg) as s:
        t0 = time.time(); s.solveAtSobolPoints(wait=False); torch.cuda.synchronize(); first = time.time() - t0
        ts = []
        for it in range(6):
            torch.cuda.synchronize(); s.solveAtSobolPoints(wait=False); torch.cuda.synchronize(); ts.append(s.stage_ms("pretest_eval"))
        ms = float(np.median(ts[1:]))
        print(name, "nx", nx, "4e6 points: kernel_ms %.4f" % ms, "%.3g evals/s" % (4e6 / (ms * 1e-3)), "first call %.2f s" % first,
              "rtc kernels", lib.load().tiktak_cuda_debug_rtc_kernels(), "NO_RTC" if os.environ.get("TIKTAK_NO_RTC") else "", flush=True)
