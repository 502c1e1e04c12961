"""GPU tests: BASELINE.json's configurations 3, 4 and 5 AT FULL SIZE against the CPU oracle (C2 at full size is in
test_gpu_parity.py::test_full_size_c2_against_oracle).

The oracle work is shared between forked processes by Sobol index range (the oracle seeks into the sequence by the Gray
code, tests/test_oracle_kat.py pins that against sequential generation) and by restart (orc_complete_search runs one restart
from the GPU's own searchStart row).  Bars (BASELINE.json): Sobol values and the selected index set bit for bit (the oracle
runs the same tk_cos as the device, include/tiktak_math.h); per-restart minima, from identical starts and caps, bit for bit
(tolerance stated by north_star: 1e-8 relative) including the evaluation counts.
"""
import multiprocessing as mp
import os

import numpy as np
import pytest

from oracle import oracle as orc
from oracle.oracle import MATH_TK, Oracle
from tiktak_b200 import TikTakSearch, baseline_config, smm_config

pytestmark = pytest.mark.gpu

_FV = None  # the GPU's fval array, inherited by the forked workers
_CFG = None


def _range_worker(args):
    first, count, keep = args
    o = Oracle(_CFG, MATH_TK)
    bad, legit, cf, ci = 0, 0, [], []
    step = 1 << 20
    for a in range(first, first + count, step):
        n = min(step, first + count - a)
        g, nl = o.sobolFnValRange(a, n)
        bad += int(np.count_nonzero(g != _FV[a - 1:a - 1 + n]))
        legit += nl
        if keep < n:
            part = np.argpartition(g, keep)[:keep]
            thr = g[part].max()
            part = np.nonzero(g <= thr)[0]  # keep every value tied with the keep-th smallest
        else:
            part = np.arange(n)
        cf.append(g[part]); ci.append(part + a)
    cf, ci = np.concatenate(cf), np.concatenate(ci)
    if len(cf) > keep:
        order = np.lexsort((ci, cf))[:keep]
        cf, ci = cf[order], ci[order]
    return bad, legit, cf, ci


def _restart_worker(args):
    alg, k, start = args
    o = Oracle(_CFG, MATH_TK)
    x, f, ne = o.completeSearch(alg, k, start)
    L = orc.load()
    if _CFG.text_roundtrip:  # result rows pass through the f40.20 record (completeSearch :594, re-read :815)
        f = L.orc_text4020(float(f))
        x = np.asarray([L.orc_text4020(float(v)) for v in x])
    return k, x, f, ne


def _point_worker(args):
    o = Oracle(_CFG, MATH_TK)
    return [(i, o.sobolFnValRange(i, 1)[0][0]) for i in args]


def _pool():
    return mp.get_context("fork").Pool(os.cpu_count() or 1)


def _check_restarts(alg, ks, st, rs, ev):
    with _pool() as pool:
        outs = pool.map(_restart_worker, [(alg, int(k), st[k - 1].copy()) for k in ks], chunksize=1)
    worst = 0.0
    for k, x, f, ne in outs:
        gx, gf = rs[k - 1, 4:], rs[k - 1, 3]
        worst = max(worst, abs(gf - f) / max(abs(f), 1e-300), float(np.max(np.abs(gx - x) / np.maximum(np.abs(x), 1e-300))))
        assert np.array_equal(gx, x) and gf == f and ev[k - 1] == ne, (k, gf, f, ev[k - 1], ne)
    assert worst <= 1e-8  # north_star's tolerance; the comparison above is in fact exact
    return len(outs)


def test_c3_full_size_against_oracle(built):
    """configs[2]: Rastrigin 20-D, 1e8 Sobol points, 10 001 Nelder-Mead restarts.  All 1e8 objective values, the valid count,
    the selected index set and rows; then 400 of the GPU's restarts re-run by the oracle from the GPU's searchStart rows."""
    global _FV, _CFG
    cfg = baseline_config("C3")
    cfg.round_size = 1024
    K, N = cfg.p_maxpoints, cfg.qr_ndraw
    with TikTakSearch(cfg) as s:
        _, ne, nl = s.solveAtSobolPoints()
        fv = s.sobolFnVal()
        xs = s.chooseSobol()
        idx = s.sobol_idx.copy()
        s.LocalMinimizations("a")
        st, rs, ev = s.searchStart.copy(), s.searchResults.copy(), s.evals.copy()
    assert ne == N and len(fv) == N
    _FV, _CFG = fv, cfg
    P = os.cpu_count() or 1
    per = -(-N // (4 * P))
    tasks = [(a, min(per, N - a + 1), K - 1) for a in range(1, N + 1, per)]
    with _pool() as pool:
        outs = pool.map(_range_worker, tasks, chunksize=1)
    assert sum(o[0] for o in outs) == 0, "Sobol-stage objective values differ from the oracle"
    assert sum(o[1] for o in outs) == nl
    cf, ci = np.concatenate([o[2] for o in outs]), np.concatenate([o[3] for o in outs])
    order = np.lexsort((ci, cf))[:K - 1]  # ascending fval, ties by ascending Sobol index (chooseSobol, stable mode)
    assert np.array_equal(ci[order], idx[1:].astype(np.int64)), "selected start-index set differs from the oracle"
    assert np.array_equal(cf[order], xs[1:, 0])
    o = Oracle(cfg, MATH_TK)
    for r in range(0, K - 1, 25):  # the rows themselves: the oracle's point of that index
        assert np.array_equal(o.sobolPointsRange(int(idx[1 + r]), 1)[0], xs[1 + r, 1:])
    assert xs[0, 0] == o.objFun(cfg.init)[0] and np.array_equal(xs[0, 1:], cfg.init)
    ks = np.unique(np.concatenate([[1, 2, 3, K], np.linspace(1, K, 400).astype(int)]))
    assert _check_restarts(1, ks, st, rs, ev) >= 400
    _FV = None


def test_c4_full_size_restarts_against_oracle(built):
    """configs[3]: Rosenbrock 50-D, 1e7 Sobol points, 20 001 Nelder-Mead restarts (most of them run into amoeba's
    5000-evaluation cap).  All 1e7 objective values + the selected set, and 300 restarts re-run by the oracle from the GPU's starts."""
    global _FV, _CFG
    cfg = baseline_config("C4")
    cfg.round_size = 1184
    K, N = cfg.p_maxpoints, cfg.qr_ndraw
    with TikTakSearch(cfg) as s:
        _, ne, nl = s.solveAtSobolPoints()
        fv = s.sobolFnVal()
        xs = s.chooseSobol()
        idx = s.sobol_idx.copy()
        s.LocalMinimizations("a")
        st, rs, ev = s.searchStart.copy(), s.searchResults.copy(), s.evals.copy()
    _FV, _CFG = fv, cfg
    P = os.cpu_count() or 1
    per = -(-N // (4 * P))
    with _pool() as pool:
        outs = pool.map(_range_worker, [(a, min(per, N - a + 1), K - 1) for a in range(1, N + 1, per)], chunksize=1)
    assert sum(o[0] for o in outs) == 0 and sum(o[1] for o in outs) == nl
    cf, ci = np.concatenate([o[2] for o in outs]), np.concatenate([o[3] for o in outs])
    order = np.lexsort((ci, cf))[:K - 1]
    assert np.array_equal(ci[order], idx[1:].astype(np.int64)) and np.array_equal(cf[order], xs[1:, 0])
    ks = np.unique(np.concatenate([[1, 2, K], np.linspace(1, K, 300).astype(int)]))
    assert _check_restarts(1, ks, st, rs, ev) >= 300
    assert np.median(ev) >= 5000  # most restarts run into the cap: a round is close to a fixed amount of work
    _FV = None


def test_c5_full_size_against_oracle(built):
    """configs[4] at BASELINE.md's size: synthetic SMM, 1200 moments, 7 parameters, panel 2^17 x 40, qr_ndraw 200 000,
    legitimate 100 000, 1001 DFNLS restarts.  60 Sobol points spread over the evaluated prefix (valid and invalid ones)
    and 4 complete DFNLS restarts, by the oracle from the GPU's searchStart rows -- one oracle evaluation simulates the whole
    panel (~0.2 s), so the sample is what fits in about a minute of host time."""
    global _CFG
    cfg = smm_config()
    cfg.round_size = 128
    assert (cfg.nm, cfg.qr_ndraw, cfg.legitimate, cfg.p_maxpoints) == (1200, 200_000, 100_000, 1001)
    K = cfg.p_maxpoints
    with TikTakSearch(cfg) as s:
        _, ne, nl = s.solveAtSobolPoints()
        fv = s.sobolFnVal()
        xs = s.chooseSobol()
        s.LocalMinimizations("b")
        st, rs, ev = s.searchStart.copy(), s.searchResults.copy(), s.evals.copy()
        assert not s.status.any(), "a DFNLS restart stopped where the reference calls RESCUE_H (not built on the device)"
    assert (nl == cfg.legitimate and fv[ne - 1] < cfg.fvalmax) or (ne == cfg.qr_ndraw and nl < cfg.legitimate)  # :429
    assert int(np.count_nonzero(fv < cfg.fvalmax)) == nl
    _CFG = cfg
    P = os.cpu_count() or 1
    probe = np.unique(np.concatenate([[1, 2, ne - 1, ne], np.linspace(1, ne, 56).astype(int)]))
    ks = [1, 2, 400, K]  # K and 400 > K/2 take the reduced radii of completeSearch :565-571
    with _pool() as pool:
        r_async = pool.map_async(_restart_worker, [(0, int(k), st[k - 1].copy()) for k in ks], chunksize=1)
        pts = pool.map(_point_worker, [probe[i::max(1, P - len(ks))].tolist() for i in range(max(1, P - len(ks)))], chunksize=1)
        routs = r_async.get()
    for i, f in (p for chunk in pts for p in chunk):
        assert fv[i - 1] == f, (i, fv[i - 1], f)
    for k, x, f, ne_o in routs:
        assert np.array_equal(rs[k - 1, 4:], x) and rs[k - 1, 3] == f and ev[k - 1] == ne_o, (k, rs[k - 1, 3], f, ev[k - 1], ne_o)
    assert np.array_equal(np.sort(xs[1:, 0]), xs[1:, 0]) and xs[1, 0] == fv.min()
