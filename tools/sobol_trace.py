"""experiment (needs the trace build: `make -C tiktak_b200/csrc trace`, passed as TIKTAK_CUDA_LIB): where block 0 of the Sobol
kernel spends its time -- SM cycle stamps and globaltimer stamps at the phase boundaries, beside the event time"""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tiktak_b200 import TikTakSearch, baseline_config, lib
L = lib.load()
names = ["entry", "index math", "tables in smem", "seed A", "seed B", "point 0", "point 1", "point 2", "last point", "exit"]
for w, n in (("C2", 1000), ("C2", 1000000), ("C4", 1000)):
    cfg = baseline_config(w); cfg.qr_ndraw = n; cfg.legitimate = n + 10000; cfg.maxpoints = 10
    with TikTakSearch(cfg) as s:
        for it in range(4):
            torch.cuda.synchronize()
            s.solveAtSobolPoints(wait=False); torch.cuda.synchronize()
            ms = s.stage_ms("pretest_eval")
        buf = (C.c_ulonglong * 32)()
        L.tiktak_cuda_debug_trace(buf)
        v = np.array(list(buf), dtype=np.int64).reshape(16, 2)
        print(w, n, "event time %.2f us" % (ms * 1e3))
        for i, nm in enumerate(names):
            print("   %-16s cycles +%7d  globaltimer +%6d ns" % (nm, v[i, 0] - v[0, 0], v[i, 1] - v[0, 1]))
