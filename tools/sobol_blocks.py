"""experiment (trace build: `make -C tiktak_b200/csrc trace`, passed as TIKTAK_CUDA_LIB): Gantt data of one C2 launch -- when every block starts and ends
(globaltimer), on which SM; summarised as resident blocks per SM over time"""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tiktak_b200 import TikTakSearch, baseline_config, lib
L = lib.load()
w, n = "C2", 1000000
cfg = baseline_config(w); cfg.qr_ndraw = n; cfg.legitimate = n + 10000; cfg.maxpoints = 10
with TikTakSearch(cfg) as s:
    for it in range(4):
        torch.cuda.synchronize()
        s.solveAtSobolPoints(wait=False); torch.cuda.synchronize()
        ms = s.stage_ms("pretest_eval")
    buf = (C.c_ulonglong * (3 * 8192))()
    L.tiktak_cuda_debug_blocks(buf)
    v = np.array(list(buf), dtype=np.int64).reshape(8192, 3)
    v = v[v[:, 0] > 0]
    t0 = v[:, 0].min()
    st, en, sm = (v[:, 0] - t0) / 1e3, (v[:, 1] - t0) / 1e3, v[:, 2]
    print("event time %.2f us; blocks %d; first start 0, last start %.2f us, last end %.2f us" % (ms * 1e3, len(v), st.max(), en.max()))
    print("block lifetime us: min %.2f med %.2f max %.2f" % ((en - st).min(), np.median(en - st), (en - st).max()))
    print("SMs used %d; blocks per SM min %d max %d" % (len(set(sm)), np.bincount(sm).min(), np.bincount(sm).max()))
    for T in np.arange(0, en.max() + 1, 2.0):
        res = ((st <= T) & (en > T)).sum()
        print("  t=%5.1f us resident blocks %5d (%.2f per SM), started %5d, finished %5d" % (T, res, res / 148, (st <= T).sum(), (en <= T).sum()))
    k = np.argsort(st)[:0]
    s0 = sm == sm[0]
    print("SM %d timeline:" % sm[0], sorted(zip(np.round(st[s0], 2), np.round(en[s0], 2))))
