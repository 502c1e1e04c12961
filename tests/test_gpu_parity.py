"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on the same inputs.

Bars (BASELINE.json north_star): Sobol points and selected start-index set bit-exact; objective values
<= 1e-12 relative vs the libm oracle (and bit-exact vs the tk-math oracle); per-restart local minima from
identical starts and caps <= 1e-8 relative (bit-exact expected); global minimum <= 1e-6.
"""
import numpy as np
import pytest

from oracle.oracle import MATH_LIBM, MATH_TK, Oracle
from tiktak_b200 import (OBJ_GRIEWANK, OBJ_RASTRIGIN, OBJ_ROSENBROCK, OBJ_TEMPLATE, TikTakSearch, baseline_config,
                         make_config, smm_config)

pytestmark = pytest.mark.gpu

SMALL = {
    "template4": dict(nx=4, nm=8, qr_ndraw=20, maxpoints=3, lo=-1.0, hi=1.0, bound=[[-3.0, 3.0]] * 4, legitimate=10,
                      fvalmax=50.0, init=[0.811530031, -0.185093302, 0.299819619, 0.196328895], objective_id=OBJ_TEMPLATE),
    "griewank10": dict(nx=10, nm=1, qr_ndraw=20000, maxpoints=40, lo=-400.0, hi=600.0, init=[100.0] * 10,
                       objective_id=OBJ_GRIEWANK),
    "rastrigin20": dict(nx=20, nm=1, qr_ndraw=70000, maxpoints=30, lo=-4.12, hi=5.12, init=[1.0] * 20,
                        objective_id=OBJ_RASTRIGIN),
    "rosenbrock50": dict(nx=50, nm=1, qr_ndraw=5000, maxpoints=12, lo=-5.0, hi=10.0, init=[2.5] * 50,
                         objective_id=OBJ_ROSENBROCK),
    # generic (runtime-nx) kernels: dimensions without a compiled-in specialisation
    "griewank7": dict(nx=7, nm=3, qr_ndraw=3000, maxpoints=10, lo=-40.0, hi=60.0, init=[10.0] * 7,
                      objective_id=OBJ_GRIEWANK),
    "rosenbrock33": dict(nx=33, nm=1, qr_ndraw=2000, maxpoints=6, lo=-2.0, hi=2.0, init=[0.5] * 33,
                         objective_id=OBJ_ROSENBROCK),
}


def cfg_of(name, **kw):
    d = dict(SMALL[name])
    d.update(kw)
    return make_config(**d)


@pytest.mark.parametrize("name", list(SMALL))
def test_sobol_points_bit_exact(built, name):
    cfg = cfg_of(name)
    n = min(cfg.qr_ndraw, 4096)
    with TikTakSearch(cfg) as s:
        g = s.setupSobol(1, n)
        far = s.setupSobol(cfg.qr_ndraw - 9, 10)
    o = Oracle(cfg, MATH_TK)
    ref = o.setupSobol(1, cfg.qr_ndraw)
    assert np.array_equal(g, ref[:n])
    assert np.array_equal(far, ref[-10:])


@pytest.mark.parametrize("name", list(SMALL))
def test_pretest_values(built, name):
    cfg = cfg_of(name)
    with TikTakSearch(cfg) as s:
        _, ne, nl = s.solveAtSobolPoints()
        f = s.sobolFnVal()
    ot = Oracle(cfg, MATH_TK)
    _, one, onl = ot.solveAtSobolPoints()
    assert (ne, nl) == (one, onl)
    ft = ot.sobolFnVal()
    assert np.array_equal(f, ft), f"max rel diff {np.max(np.abs(f - ft) / np.abs(ft))}"
    ol = Oracle(cfg, MATH_LIBM)
    ol.solveAtSobolPoints()
    fl = ol.sobolFnVal()
    assert np.max(np.abs(f - fl) / np.abs(fl)) <= 1e-12  # tolerance from BASELINE.json


@pytest.mark.parametrize("name,blo,bhi,fvalmax", [("template4", -0.5, 0.75, 1e4), ("griewank10", -300.0, 500.0, 1e4),
                                                  ("rastrigin20", -4.0, 5.0, 1e7), ("rosenbrock50", -4.9, 9.9, 1e18),
                                                  ("griewank7", -30.0, 50.0, 1e4)])
def test_pretest_values_box_outside_bounds(built, name, blo, bhi, fvalmax):
    """Sobol box wider than the bounds: the getPenalty branch of dfovec / objFun (objective.f90:51-74, 92-95) inside the
    Sobol kernel (the bounds test is compiled out only when the host has verified that the box lies inside the bounds)"""
    nx = SMALL[name]["nx"]
    cfg = cfg_of(name, bound=[[blo, bhi]] * nx, fvalmax=fvalmax, legitimate=-1)
    with TikTakSearch(cfg) as s:
        _, ne, nl = s.solveAtSobolPoints()
        f = s.sobolFnVal()
    o = Oracle(cfg, MATH_TK)
    _, one, onl = o.solveAtSobolPoints()
    fo = o.sobolFnVal()
    assert (ne, nl) == (one, onl)
    assert np.array_equal(f, fo)
    assert 0 < np.count_nonzero(f >= cfg.fvalmax) < len(f), "the case must mix penalised and regular points"


def test_pretest_values_outside_trig_domain(built):
    """arguments beyond tk_cos's domain (|x| >= 2^30): the plugin's domain_ok fails on the host, the kernel keeps the
    per-call domain test, and device and oracle agree on which values are NaN and on every other value"""
    cfg = cfg_of("griewank10", lo=-3.0e9, hi=3.0e9, qr_ndraw=6000, init=[1.0] * 10)
    with TikTakSearch(cfg) as s:
        _, ne, nl = s.solveAtSobolPoints()
        f = s.sobolFnVal()
    o = Oracle(cfg, MATH_TK)
    _, one, onl = o.solveAtSobolPoints()
    fo = o.sobolFnVal()
    assert (ne, nl) == (one, onl)
    assert np.array_equal(np.isnan(f), np.isnan(fo))
    ok = ~np.isnan(fo)
    assert 0 < np.count_nonzero(ok) < len(fo)
    assert np.array_equal(f[ok], fo[ok])


@pytest.mark.parametrize("name", ["griewank10", "rastrigin20", "griewank7"])
def test_legitimate_cut(built, name):
    """finite fvalmax and legitimate < #valid: the evaluated set is the prefix 1..Kcut (:429-464)"""
    base = cfg_of(name)
    o0 = Oracle(base, MATH_TK)
    o0.solveAtSobolPoints()
    med = float(np.median(o0.sobolFnVal()))
    for L in (1, 17, base.qr_ndraw // 4):
        cfg = cfg_of(name, fvalmax=med, legitimate=L)
        with TikTakSearch(cfg) as s:
            _, ne, nl = s.solveAtSobolPoints()
            f = s.sobolFnVal()
            xs = s.chooseSobol()
            idx = s.sobol_idx.copy()
        o = Oracle(cfg, MATH_TK)
        _, one, onl = o.solveAtSobolPoints()
        assert (ne, nl) == (one, onl), (L, ne, nl, one, onl)
        assert np.array_equal(f, o.sobolFnVal())
        if ne >= cfg.maxpoints:
            oxs = o.chooseSobol(mode=1)
            assert np.array_equal(idx, o.sobol_idx)
            assert np.array_equal(xs, oxs)


@pytest.mark.parametrize("name", list(SMALL))
def test_choose_sobol(built, name):
    cfg = cfg_of(name)
    with TikTakSearch(cfg) as s:
        s.solveAtSobolPoints()
        xs = s.chooseSobol()
        idx = s.sobol_idx.copy()
    o = Oracle(cfg, MATH_TK)
    o.solveAtSobolPoints()
    oxs = o.chooseSobol(mode=1)
    assert np.array_equal(idx, o.sobol_idx)
    assert np.array_equal(xs, oxs)
    # against the reference's own (unstable) indexx order: same SET unless equal keys straddle the cut
    oxs0 = o.chooseSobol(mode=0)
    if not o.tie_straddle:
        assert set(idx[1:].tolist()) == set(o.sobol_idx[1:].tolist())


@pytest.mark.parametrize("name,B", [("template4", 1), ("template4", 4), ("griewank10", 1), ("griewank10", 8),
                                    ("rastrigin20", 6), ("rosenbrock50", 13), ("griewank7", 3), ("rosenbrock33", 7)])
def test_local_minimizations_nm(built, name, B):
    cfg = cfg_of(name, round_size=B)
    with TikTakSearch(cfg) as s:
        s.solveAtSobolPoints()
        s.chooseSobol()
        s.LocalMinimizations("a")
        gs, gr, ge = s.searchStart, s.searchResults, s.evals
        gbest = s.getBestLine()
    o = Oracle(cfg, MATH_TK)
    o.solveAtSobolPoints()
    o.chooseSobol(mode=1)
    o.LocalMinimizations(1)
    assert np.array_equal(gs, o.searchStart), "blended start points differ"
    rel = np.abs(gr[:, 3] - o.searchResults[:, 3]) / np.abs(o.searchResults[:, 3])
    assert np.max(rel) <= 1e-8, rel  # BASELINE.json bar; bit-exact expected:
    assert np.array_equal(gr, o.searchResults)
    assert np.array_equal(ge, o.evals)
    ob = o.getBestLine()
    assert abs(gbest[0] - ob[0]) <= 1e-6 * abs(ob[0]) and gbest[2] == ob[2]


@pytest.mark.parametrize("name,P,nm,K", [("griewank10", 8, 1, 40), ("griewank10", 16, 3, 40), ("griewank10", 64, 1, 200), ("rastrigin20", 7, 1, 30),
                                          ("rastrigin20", 30, 2, 30), ("template4", 3, 8, 3), ("griewank7", 4, 3, 10), ("griewank7", 1, 1, 10)])
def test_asynchronous_instances(built, name, P, nm, K):
    """the reference's multi-process mode on one GPU (tiktak_cuda_local_async, nm_quad_async_kernel): P blocks that take the
    next restart the moment they finish one, every start blended toward the best row finished by then, reproducible through a
    virtual clock -- against the oracle's sequential event simulation of P instances (Oracle::local_async): start rows, result
    rows incl. the #improvements column in file order, evaluation counts and the best row, bit for bit"""
    cfg = cfg_of(name, nm=nm, maxpoints=K, round_size=P)
    with TikTakSearch(cfg) as s:
        assert s.asyncSupported("a")
        s.solveAtSobolPoints()
        s.chooseSobol()
        s.LocalMinimizationsAsync("a")
        gs, gr, ge, used, clock = s.searchStart, s.searchResults, s.evals, s.instances_used, s.asyncClock("a")
        gbest = s.getBestLine()
    assert used == min(P, cfg.p_maxpoints)
    o = Oracle(cfg, MATH_TK)
    o.solveAtSobolPoints()
    o.chooseSobol(mode=1)
    o.LocalMinimizationsAsync(used, clock=clock)
    assert np.array_equal(gs, o.searchStart), "blended start points differ"
    assert np.array_equal(ge, o.evals)
    assert np.array_equal(gr, o.searchResults)
    ob = o.getBestLine()
    assert gbest[0] == ob[0] and gbest[2] == ob[2]
    if P > 1:  # the instances did overlap: some restart was blended toward a row other than the best of all earlier restarts
        assert len(set(o.inst.tolist())) == used
        if cfg.p_maxpoints > 2 * used:
            assert np.any(np.diff(o.fin) < 0)


@pytest.mark.parametrize("name,P,nm,K,mode", [("griewank10", 16, 3, 40, None), ("griewank10", 64, 1, 200, None), ("rastrigin20", 30, 2, 30, None),
                                               ("rastrigin20", 200, 1, 600, None), ("template4", 3, 8, 3, None), ("griewank7", 1, 1, 10, None),
                                               ("rosenbrock33", 5, 1, 12, None), ("rosenbrock50", 40, 1, 90, None), ("rastrigin20", 30, 1, 60, "324"),
                                               ("griewank10", 20, 1, 50, "81")])
def test_asynchronous_instances_one_warp(built, monkeypatch, name, P, nm, K, mode):
    """the throughput form of the asynchronous run: ONE WARP is the instance (nm_w4_kernel with NmQueue.aq; nx 32..127 by default,
    TIKTAK_ASYNC_FORM=w4 below that), the tick of the virtual clock is one objective evaluation (TIKTAK_ASYNC_CLOCK_EVALS) --
    against the oracle's event simulation on that clock, bit for bit"""
    monkeypatch.setenv("TIKTAK_ASYNC_FORM", "w4")
    if mode:
        monkeypatch.setenv("TIKTAK_ASYNC_MODE", mode)
    cfg = cfg_of(name, nm=nm, maxpoints=K, round_size=P)
    with TikTakSearch(cfg) as s:
        assert s.asyncSupported("a") and s.asyncClock("a") == 1
        s.solveAtSobolPoints()
        s.chooseSobol()
        s.LocalMinimizationsAsync("a")
        gs, gr, ge, used = s.searchStart, s.searchResults, s.evals, s.instances_used
        gbest = s.getBestLine()
    assert used == min(P, cfg.p_maxpoints)
    o = Oracle(cfg, MATH_TK)
    o.solveAtSobolPoints()
    o.chooseSobol(mode=1)
    o.LocalMinimizationsAsync(used, clock=1)
    assert np.array_equal(gs, o.searchStart), "blended start points differ"
    assert np.array_equal(ge, o.evals)
    assert np.array_equal(gr, o.searchResults)
    ob = o.getBestLine()
    assert gbest[0] == ob[0] and gbest[2] == ob[2]
    if used > 1:
        assert len(set(o.inst.tolist())) == used


@pytest.mark.parametrize("name,P,nm,K", [("template4", 3, 8, 9), ("griewank7", 4, 3, 10), ("griewank10", 16, 6, 40), ("rosenbrock33", 5, 4, 12)])
def test_asynchronous_instances_dfnls(built, name, P, nm, K):
    """the same scheme for DFNLS (dfnls_kernel with DfArgs.async): the tick of the virtual clock is one objective evaluation"""
    cfg = cfg_of(name, nm=nm, maxpoints=K, round_size=P)
    with TikTakSearch(cfg) as s:
        assert s.asyncSupported("b")
        s.solveAtSobolPoints()
        s.chooseSobol()
        s.LocalMinimizationsAsync("b")
        gs, gr, ge, used = s.searchStart, s.searchResults, s.evals, s.instances_used
        gbest = s.getBestLine()
    assert used == min(P, cfg.p_maxpoints)
    o = Oracle(cfg, MATH_TK)
    o.solveAtSobolPoints()
    o.chooseSobol(mode=1)
    o.LocalMinimizationsAsync(used, algor=0)
    assert np.array_equal(gs, o.searchStart), "blended start points differ"
    assert np.array_equal(ge, o.evals)
    assert np.array_equal(gr, o.searchResults)
    assert gbest[0] == o.getBestLine()[0]


def test_asynchronous_instances_smm_dfnls(built):
    """BASELINE config 5 at reduced panel size, DFNLS by asynchronous instances (the form bench.py times at full size)"""
    cfg = smm_config(np_persons=1024, T=20, qr_ndraw=300, legitimate=40, maxpoints=11, fvalmax=20000.0, round_size=4)
    with TikTakSearch(cfg) as s:
        s.solveAtSobolPoints()
        s.chooseSobol()
        s.LocalMinimizationsAsync("b")
        gs, gr, ge, used = s.searchStart, s.searchResults, s.evals, s.instances_used
    o = Oracle(cfg, MATH_TK)
    o.solveAtSobolPoints()
    o.chooseSobol(mode=1)
    o.LocalMinimizationsAsync(used, algor=0)
    assert used == 4 and np.array_equal(gs, o.searchStart) and np.array_equal(ge, o.evals) and np.array_equal(gr, o.searchResults)


@pytest.mark.parametrize("form", ["quad", "w4"])
def test_free_running_instances_restart_by_restart(built, monkeypatch, form):
    """(as four-warp blocks and as one-warp instances)  instances < 0: no virtual clock -- like the reference's processes, an instance blends toward whatever row is the best one
    committed when it takes its next restart, so the run is not reproducible as a whole.  Checked restart by restart: every restart
    ran once; its start is the blend of its Sobol start with the row the library says it used (a row committed before it, or none);
    and the search from that start is the oracle's, bit for bit (minimum, value, evaluation count)."""
    import ctypes as C
    from tiktak_b200._abi import dptr
    monkeypatch.setenv("TIKTAK_ASYNC_FORM", form)
    cfg = cfg_of("griewank10", maxpoints=150, round_size=48, text_roundtrip=0)  # (rows as computed, not through the f40.20 record)
    K, nx = cfg.p_maxpoints, cfg.nx
    with TikTakSearch(cfg) as s:
        s.solveAtSobolPoints()
        xs = s.chooseSobol()
        s.LocalMinimizationsAsync("a", free_running=True)
        st, res, ev, base, order, used = s.searchStart, s.searchResults, s.evals, s.async_base, s.async_order, s.instances_used
    assert used == 48 and sorted(order.tolist()) == list(range(1, K + 1))
    o = Oracle(cfg, MATH_TK)
    o.solveAtSobolPoints()
    assert np.array_equal(xs, o.chooseSobol(mode=1))
    blended = 0
    for k in range(1, K + 1):
        b = int(base[k - 1])
        if b < 0:
            want = xs[k - 1, 1:]
        else:
            assert b == 0 or order[b - 1] < order[k - 1]
            w = np.float64(min(max(np.float32(0.10), np.sqrt(np.float32(k) / np.float32(K))), np.float32(0.95)))
            bx = xs[0, 1:] if b == 0 else res[b - 1, 4:]
            want = w * bx + (1.0 - w) * xs[k - 1, 1:]
            blended += 1
        assert np.array_equal(st[k - 1], want), k
        pt = np.ascontiguousarray(st[k - 1].copy())
        it = C.c_int32()
        assert o.L.orc_run_amoeba(o.h, dptr(pt), C.byref(it)) == 0
        assert np.array_equal(pt, res[k - 1, 4:]) and o.objFun(pt)[0] == res[k - 1, 3] and ev[k - 1] == nx + 1 + it.value + 1, k
    assert blended >= K - 48  # everything after the first pick-up of each instance saw at least one finished row


def test_one_asynchronous_instance_is_the_sequential_run(built):
    """P = 1: the instance sees every earlier row -- the reference's single-process run, i.e. rounds of one restart"""
    cfg = cfg_of("griewank10", maxpoints=25, round_size=1)
    with TikTakSearch(cfg) as s:
        s.solveAtSobolPoints()
        s.chooseSobol()
        s.LocalMinimizationsAsync("a", instances=1)
        a_s, a_r, a_e = s.searchStart, s.searchResults, s.evals
    with TikTakSearch(cfg) as s:
        s.solveAtSobolPoints()
        s.chooseSobol()
        s.LocalMinimizations("a")
        assert np.array_equal(a_s, s.searchStart) and np.array_equal(a_r, s.searchResults) and np.array_equal(a_e, s.evals)


@pytest.mark.parametrize("name,B,nm", [("template4", 1, 8), ("template4", 4, 8), ("griewank10", 5, 6), ("griewank7", 4, 3),
                                       ("rastrigin20", 8, 2), ("rosenbrock33", 3, 4)])
def test_local_minimizations_dfnls(built, name, B, nm):
    """DFNLS (bobyqa_h family, minimize.f90:262-2563), one restart per block, vs the serial restatement"""
    cfg = cfg_of(name, round_size=B, nm=nm, maxpoints=min(SMALL[name]["maxpoints"], 9))
    with TikTakSearch(cfg) as s:
        s.solveAtSobolPoints()
        s.chooseSobol()
        s.LocalMinimizations("b")
        gs, gr, ge = s.searchStart, s.searchResults, s.evals
    o = Oracle(cfg, MATH_TK)
    o.solveAtSobolPoints()
    o.chooseSobol(mode=1)
    o.LocalMinimizations(0)
    assert np.array_equal(gs, o.searchStart)
    rel = np.abs(gr[:, 3] - o.searchResults[:, 3]) / np.abs(o.searchResults[:, 3])
    assert np.max(rel) <= 1e-8, rel  # BASELINE.json bar; bit-exact expected:
    assert np.array_equal(ge, o.evals)
    assert np.array_equal(gr, o.searchResults)


@pytest.mark.parametrize("name,B,nm,first,every", [("template4", 4, 8, 11, 0), ("griewank7", 4, 3, 17, 0), ("griewank7", 4, 3, 26, 7),
                                                   ("griewank10", 5, 6, 30, 9), ("rastrigin20", 8, 2, 45, 40), ("rosenbrock33", 3, 4, 70, 0)])
def test_dfnls_rescue_forced(built, name, B, nm, first, every, monkeypatch):
    """RESCUE_H on the device (minimize.f90:1627-2023, label 190 :643-675).  Rounding damages a denominator in almost no run
    (0 entries in 3600 oracle runs over the four objectives, 0 at C5), so the test hook makes the denominator tests of :746 /
    :782 count as failed at chosen evaluation numbers, in the kernel (TIKTAK_DFNLS_FORCE_RESCUE) and in the oracle
    (orc_set_dfnls_force_rescue) alike.  Both must then walk RESCUE_H through the same exchanges of provisional and original
    points, make the same evaluations inside it and end every restart on the same row, bit for bit."""
    cfg = cfg_of(name, round_size=B, nm=nm, maxpoints=min(SMALL[name]["maxpoints"], 9))
    monkeypatch.setenv("TIKTAK_DFNLS_FORCE_RESCUE", "%d,%d" % (first, every))
    with TikTakSearch(cfg) as s:
        s.solveAtSobolPoints()
        s.chooseSobol()
        s.LocalMinimizations("b")
        gs, gr, ge, gn = s.searchStart, s.searchResults, s.evals, s.nrescue
        assert np.all(s.status == 0)
    o = Oracle(cfg, MATH_TK)
    o.L.orc_set_dfnls_force_rescue(o.h, first, every)
    o.solveAtSobolPoints()
    o.chooseSobol(mode=1)
    o.LocalMinimizations(0)
    n_resc = o.L.orc_set_dfnls_force_rescue(o.h, 0, 0)
    assert n_resc >= cfg.p_maxpoints - 1 and int(gn.sum()) == n_resc, (gn, n_resc)  # (a restart may end before `first`)
    assert np.array_equal(gs, o.searchStart)
    assert np.array_equal(ge, o.evals)
    assert np.array_equal(gr, o.searchResults)
    # and the hook switched off gives the plain run again (the knob is read at every launch)
    monkeypatch.delenv("TIKTAK_DFNLS_FORCE_RESCUE")
    with TikTakSearch(cfg) as s:
        s.solveAtSobolPoints()
        s.chooseSobol()
        s.LocalMinimizations("b")
        assert int(s.nrescue.sum()) == 0


@pytest.mark.parametrize("name,B,nm,K,first,every", [("griewank7", 5, 3, 9, 24, 2), ("rosenbrock33", 3, 4, 4, 70, 2)])
def test_dfnls_rescue_forced_with_evaluations_inside(built, name, B, nm, K, first, every, monkeypatch):
    """the second half of RESCUE_H (:1892-2019): provisional points that could not be exchanged for original ones stay, each
    costs an evaluation and a model update -- the forced schedules below (a rescue every other evaluation) leave dozens in
    (the oracle counts them); rows, evaluation counts and the number of rescues per job must agree"""
    cfg = cfg_of(name, round_size=B, nm=nm, maxpoints=K)
    monkeypatch.setenv("TIKTAK_DFNLS_FORCE_RESCUE", "%d,%d" % (first, every))
    with TikTakSearch(cfg) as s:
        s.solveAtSobolPoints()
        s.chooseSobol()
        s.LocalMinimizations("b")
        gr, ge, gn = s.searchResults, s.evals, s.nrescue
    o = Oracle(cfg, MATH_TK)
    o.L.orc_set_dfnls_force_rescue(o.h, first, every)
    o.solveAtSobolPoints()
    o.chooseSobol(mode=1)
    o.LocalMinimizations(0)
    assert o.L.orc_dfnls_rescue_evals(o.h) > 0
    n_resc = o.L.orc_set_dfnls_force_rescue(o.h, 0, 0)
    assert np.array_equal(ge, o.evals) and np.array_equal(gr, o.searchResults)
    assert int(gn.sum()) == n_resc or np.any(gn == 31)  # the per-restart counter saturates at 31


@pytest.mark.parametrize("nx,nm", [(64, 2), (100, 1)])
def test_dfnls_up_to_nx_100(built, nx, nm):
    """nmax = 100 of minimize.f90:416: above nx ~ 55 XPT, BMAT and ZMAT (560 KB at nx = 100) live in the restart's slab of
    global memory instead of shared memory; same arithmetic, same rows"""
    cfg = make_config(nx=nx, nm=nm, qr_ndraw=400, maxpoints=3, lo=-2.0, hi=2.0, init=[0.5] * nx, objective_id=OBJ_ROSENBROCK,
                      round_size=2, maxeval=3 * nx + 40)
    with TikTakSearch(cfg) as s:
        s.solveAtSobolPoints()
        s.chooseSobol()
        s.LocalMinimizations("b")
        gs, gr, ge = s.searchStart, s.searchResults, s.evals
        assert np.all(s.status == 0)
    o = Oracle(cfg, MATH_TK)
    o.solveAtSobolPoints()
    o.chooseSobol(mode=1)
    o.LocalMinimizations(0)
    assert np.array_equal(gs, o.searchStart)
    assert np.array_equal(ge, o.evals)
    assert np.array_equal(gr, o.searchResults)


def test_smm_objective_pretest_and_dfnls(built):
    """BASELINE config 5 at reduced panel size: synthetic SMM (30 T moments, 7 parameters, panel simulated per
    evaluation), legitimate cut with a finite fvalmax, DFNLS local stage."""
    cfg = smm_config(np_persons=1024, T=20, qr_ndraw=300, legitimate=40, maxpoints=5, fvalmax=20000.0, round_size=3)
    o = Oracle(cfg, MATH_TK)
    with TikTakSearch(cfg) as s:
        pts = np.vstack([cfg.extra["smm"]["theta_star"], cfg.init, [0.999, 0.1, 0.1, 0.1, 0.1, 0.5, 0.0], [0.7, 2.0, 0.3, 0.3, 0.3, 0.5, 0.1]])
        fg = s.objFun(pts)
        assert np.array_equal(fg, o.objFun(pts))  # incl. one out-of-bounds point (penalty branch)
        _, ne, nl = s.solveAtSobolPoints()
        _, one, onl = o.solveAtSobolPoints()
        assert (ne, nl) == (one, onl) and nl == 40 and ne < 300
        assert np.array_equal(s.sobolFnVal(), o.sobolFnVal())
        xs = s.chooseSobol()
        assert np.array_equal(xs, o.chooseSobol(mode=1))
        s.LocalMinimizations("b")
        o.LocalMinimizations(0)
        assert np.array_equal(s.searchStart, o.searchStart)
        assert np.array_equal(s.evals, o.evals)
        assert np.array_equal(s.searchResults, o.searchResults)
        fb, xb, kb = s.getBestLine()
    assert fb < 5.0 and abs(xb[0] - 0.9) < 0.1  # rho recovered near theta_star


@pytest.mark.parametrize("name,B", [("griewank10", 5), ("rosenbrock50", 4)])
def test_nm_shrink_mode(built, name, B):
    """nm_shrink_mode = 1 (README.md:36 rule; SURVEY finding 3): later restarts scale their initial simplex by the
    spread of the best minima so far.  Same rule in the oracle -> bit-exact; and it must change something."""
    cfg = cfg_of(name, round_size=B, nm_shrink_mode=1)
    with TikTakSearch(cfg) as s:
        s.solveAtSobolPoints(); s.chooseSobol(); s.LocalMinimizations("a")
        gr, ge = s.searchResults, s.evals
    o = Oracle(cfg, MATH_TK)
    o.solveAtSobolPoints(); o.chooseSobol(mode=1); o.LocalMinimizations(1)
    assert np.array_equal(gr, o.searchResults) and np.array_equal(ge, o.evals)
    o0 = Oracle(cfg_of(name, round_size=B, nm_shrink_mode=0), MATH_TK)
    o0.solveAtSobolPoints(); o0.chooseSobol(mode=1); o0.LocalMinimizations(1)
    assert not np.array_equal(o0.searchResults, o.searchResults)


@pytest.mark.parametrize("name,alg", [("template4", "a"), ("griewank10", "a"), ("rosenbrock33", "a"), ("griewank7", "b")])
def test_last_search_bfgs(built, name, alg):
    """state 10: EST_dfpmin polish from the best row (minimize.f90:2572-2816); final global minimum <= 1e-6"""
    cfg = cfg_of(name, round_size=4)
    with TikTakSearch(cfg) as s:
        s.solveAtSobolPoints(); s.chooseSobol(); s.LocalMinimizations(alg)
        fb = s.getBestLine()[0]
        f, x, it = s.lastSearch()
    o = Oracle(cfg, MATH_TK)
    o.solveAtSobolPoints(); o.chooseSobol(mode=1); o.LocalMinimizations(1 if alg == "a" else 0)
    of, ox, oit = o.lastSearch()
    assert abs(f - of) <= 1e-6 * abs(of)  # BASELINE.json bar
    assert f == of and np.array_equal(x, ox) and it == oit  # and in fact bit-exact
    assert f <= fb


def test_cli_driver_writes_reference_files(built, tmp_path):
    """driver/TiktakGlobalSearch 0 config a: same CLI, config.txt grammar and .dat record formats as the reference"""
    import os, subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "driver", "TiktakGlobalSearch")
    cfgfile = os.path.join(root, "tests", "golden", "config_c1.txt")
    r = subprocess.run([exe, "0", cfgfile, "a"], cwd=tmp_path, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert int(open(tmp_path / "state.dat").read()) == 10
    rows = np.loadtxt(tmp_path / "searchResults.dat")
    fn = np.loadtxt(tmp_path / "sobolFnVal.dat")
    xs = np.loadtxt(tmp_path / "x_starts.dat")
    fin = np.loadtxt(tmp_path / "FinalResults.dat")
    line = open(tmp_path / "searchResults.dat").readline().rstrip("\n")
    assert len(line) == 3 * 10 + 5 * 40  # (3i10, 200f40.20) with nx = 4
    assert len(open(tmp_path / "sobolFnVal.dat").readline().rstrip("\n")) == 8 + 5 * 40  # (i8, 200f40.20)
    cfg = baseline_config("C1")
    o = Oracle(cfg, MATH_TK)
    o.solveAtSobolPoints(); oxs = o.chooseSobol(mode=1); o.LocalMinimizations(1)
    assert fn.shape == (10, 6) and np.array_equal(fn[:, 0], np.arange(1, 11))
    assert np.allclose(fn[:, 1], o.sobolFnVal(), rtol=1e-15, atol=0)
    assert np.allclose(xs, oxs, rtol=0, atol=1e-15)  # F30.15
    assert rows.shape == (5, 8) and rows[0, 1] == 0  # the k = 0 record first (:496-501)
    assert np.allclose(rows[1:], o.searchResults, rtol=1e-15, atol=1e-19)
    of, ox, _ = o.lastSearch()
    assert abs(fin[2] - of) <= 1e-6 * abs(of)
    assert int(open(tmp_path / "lastSobol.dat").read()) == 11 and int(open(tmp_path / "legitSobol.dat").read()) == 10


def test_cli_driver_asynchronous_instances(built, tmp_path, monkeypatch):
    """TIKTAK_ASYNC_INSTANCES=3: the CLI driver runs state 7 through tiktak_cuda_local_async; rows as the oracle's event simulation"""
    import os, subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "driver", "TiktakGlobalSearch")
    cfgfile = os.path.join(root, "tests", "golden", "config_c1.txt")
    monkeypatch.setenv("TIKTAK_ASYNC_INSTANCES", "3")
    r = subprocess.run([exe, "0", cfgfile, "a"], cwd=tmp_path, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    rows = np.loadtxt(tmp_path / "searchResults.dat")
    o = Oracle(baseline_config("C1"), MATH_TK)
    o.solveAtSobolPoints(); o.chooseSobol(mode=1); o.LocalMinimizationsAsync(3)
    assert rows.shape == (5, 8) and rows[0, 1] == 0
    assert np.allclose(rows[1:], o.searchResults, rtol=1e-15, atol=1e-19)
    assert "as 3 asynchronous instances" in open(tmp_path / "logFile.txt").read()


def test_run_time_compiled_sobol_kernel_is_used_and_agrees_with_the_runtime_nx_kernel(built, tmp_path):
    """(objective, nx) pairs without a compiled-in instantiation get the fixed-nx kernel through NVRTC (rtc.cu); TIKTAK_NO_RTC=1
    keeps the runtime-nx kernel.  Both against the oracle, bit for bit, with the Sobol box inside and outside the bounds."""
    import ctypes as C, os, subprocess, sys
    from tiktak_b200 import lib
    L = lib.load()
    for name, kw in (("griewank7", {}), ("rosenbrock33", {}), ("griewank7", dict(bound=[[-30.0, 50.0]] * 7, fvalmax=1e4, legitimate=-1))):
        cfg = cfg_of(name, **kw)
        with TikTakSearch(cfg) as s:
            s.solveAtSobolPoints()
            f = s.sobolFnVal()
        o = Oracle(cfg, MATH_TK)
        o.solveAtSobolPoints()
        assert np.array_equal(f, o.sobolFnVal()), name
    assert L.tiktak_cuda_debug_rtc_kernels() >= 3  # (cached per process: earlier tests may have compiled the same kernels)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = ("import sys, numpy as np; sys.path.insert(0, %r); sys.path.insert(0, %r + '/tests')\n"
            "from tiktak_b200 import TikTakSearch, lib\nfrom test_gpu_parity import cfg_of\n"
            "s = TikTakSearch(cfg_of('griewank7')); s.solveAtSobolPoints(); np.save(%r, s.sobolFnVal()); "
            "print('rtc kernels', lib.load().tiktak_cuda_debug_rtc_kernels())\n") % (root, root, str(tmp_path / "f.npy"))
    r = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, TIKTAK_NO_RTC="1"), capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "rtc kernels 0" in r.stdout, r.stdout + r.stderr
    o = Oracle(cfg_of("griewank7"), MATH_TK)
    o.solveAtSobolPoints()
    assert np.array_equal(np.load(tmp_path / "f.npy"), o.sobolFnVal())


def test_repeated_jobs_on_one_handle_through_the_stage_graph(built):
    """the one-wave pre-test stage is captured as a CUDA graph on its second and third run (the two counts buffers alternate)
    and launched as a graph from the fourth on: values, counts and the selection stay those of the first run / the oracle"""
    cfg = cfg_of("griewank10")
    o = Oracle(cfg, MATH_TK)
    _, one, onl = o.solveAtSobolPoints()
    oxs = o.chooseSobol(mode=1)
    with TikTakSearch(cfg) as s:
        for rep in range(6):
            _, ne, nl = s.solveAtSobolPoints(wait=(rep % 2 == 0))
            assert ne == one and s.n_legit == onl
            assert np.array_equal(s.sobolFnVal(), o.sobolFnVal())
            assert np.array_equal(s.chooseSobol(), oxs)


def test_start_selection_large_m(built):
    """more winners than one sort chunk (1024): chunk sort + merge-by-rank passes, odd and even pass counts"""
    for maxpoints, nd in ((1500, 4000), (3000, 20000), (5000, 6000)):
        cfg = cfg_of("griewank10", qr_ndraw=nd, maxpoints=maxpoints)
        with TikTakSearch(cfg) as s:
            s.solveAtSobolPoints(wait=False)
            assert s._n_legit is None  # queued only; the valid count is read back on first use
            xs = s.chooseSobol()
            idx = s.sobol_idx.copy()
            nl = s.n_legit
        o = Oracle(cfg, MATH_TK)
        _, _, onl = o.solveAtSobolPoints()
        oxs = o.chooseSobol(mode=1)
        assert nl == onl
        assert np.array_equal(idx, o.sobol_idx)
        assert np.array_equal(xs, oxs)


@pytest.mark.parametrize("lo,hi,nd,maxpoints", [(-0.01, 0.01, 600_000, 50), (-400.0, 600.0, 300_000, 700), (-1e-9, 1e-9, 40_000, 30)])
def test_start_selection_threshold_list(built, lo, hi, nd, maxpoints):
    """the radix select's short list of elements that share the threshold's top 16 key bits: a narrow box puts EVERY value
    into one such bucket (the list overflows its capacity of an eighth of the points and the later passes fall back to
    the array), the C2 box leaves a small bucket, and a box of width 2e-9 makes most values exactly equal (ties at the
    threshold resolved by Sobol index through the list)"""
    cfg = cfg_of("griewank10", lo=lo, hi=hi, qr_ndraw=nd, maxpoints=maxpoints, init=[0.5 * (lo + hi)] * 10)
    with TikTakSearch(cfg) as s:
        s.solveAtSobolPoints()
        xs = s.chooseSobol()
        idx = s.sobol_idx.copy()
    o = Oracle(cfg, MATH_TK)
    o.solveAtSobolPoints()
    oxs = o.chooseSobol(mode=1)
    assert np.array_equal(idx, o.sobol_idx)
    assert np.array_equal(xs, oxs)


def test_blocking_round_api_matches_queued_rounds(built):
    """tiktak_cuda_local (one blocking call per round) and the queued enqueue/ingest/fetch form give the same rows"""
    import ctypes as C
    from tiktak_b200 import lib
    cfg = cfg_of("griewank10", round_size=7)
    L = lib.load()
    with TikTakSearch(cfg) as s:
        s.solveAtSobolPoints(); s.chooseSobol(); s.LocalMinimizations("a")
        qs, qr, qe = s.searchStart.copy(), s.searchResults.copy(), s.evals.copy()
    with TikTakSearch(cfg) as s:
        s.solveAtSobolPoints(); s.chooseSobol()
        K, nx = s.K, cfg.nx
        bs, br, be = np.empty((K, nx)), np.empty((K, nx + 4)), np.zeros(K, dtype=np.int32)
        for first_k, n_k in s.rounds():
            st = np.empty((nx, n_k)); r = np.empty((nx + 4, n_k)); e = np.zeros(n_k, dtype=np.int32)
            lib.check(L.tiktak_cuda_local(s._h, 1, first_k, n_k, st.ctypes.data, r.ctypes.data, e.ctypes.data), "local")
            sl = slice(first_k - 1, first_k - 1 + n_k)
            bs[sl], br[sl], be[sl] = st.T, r.T, e
        assert s.stage_ms("local") > 0.0
    assert np.array_equal(qs, bs) and np.array_equal(qr, br) and np.array_equal(qe, be)


@pytest.mark.parametrize("name,alg,world", [("griewank10", "a", 2), ("rastrigin20", "a", 3), ("griewank7", "b", 2)])
def test_sharded_protocol_on_one_gpu(built, name, alg, world):
    """The whole N>1 protocol without NCCL: `world` handles (rank r of world) on the same GPU; the all-gathers are
    replaced by concatenating the ranks' device buffers.  Exercises the block-cyclic Sobol shards, the device-side
    candidate export/merge, pack -> gather -> ingest of every restart round, and the shard fval copy."""
    import ctypes as C
    import torch
    from tiktak_b200 import SOBOL_BLOCK, lib
    from tiktak_b200 import dist as D
    kw = dict(qr_ndraw=3 * SOBOL_BLOCK + 4321, maxpoints=24, round_size=5)
    if alg == "b":
        kw.update(nm=3, maxpoints=8, round_size=3)
    cfg = cfg_of(name, **kw)
    L = lib.load()
    K, nx, N = cfg.p_maxpoints, cfg.nx, cfg.qr_ndraw
    nblocks = N // SOBOL_BLOCK + 1
    dev = torch.device("cuda", 0)
    o = Oracle(cfg, MATH_TK)
    o.solveAtSobolPoints(); oxs = o.chooseSobol(mode=1); o.LocalMinimizations(1 if alg == "a" else 0)
    hs = [TikTakSearch(cfg, device=0, rank=r, world=world, comm=False) for r in range(world)]
    try:
        # --- Sobol stage: every shard evaluates its own count blocks, no exchange (legitimate > qr_ndraw)
        for s in hs:
            lib.check(L.tiktak_cuda_pretest_wave(s._h, 0, nblocks), "wave")
            lib.check(L.tiktak_cuda_pretest_set_cut(s._h, N, -1), "set_cut")
            s.n_evaluated = N
        fo = o.sobolFnVal()
        for r, s in enumerate(hs):
            loc = s.sobolFnValShard()
            for q in range(len(loc) // SOBOL_BLOCK):
                b = r + q * world
                lo, hi = max(b * SOBOL_BLOCK, 1), min((b + 1) * SOBOL_BLOCK - 1, N)
                assert np.array_equal(loc[q * SOBOL_BLOCK + lo % SOBOL_BLOCK: q * SOBOL_BLOCK + hi % SOBOL_BLOCK + 1], fo[lo - 1: hi])
        # --- start selection: local top-(K-1) -> exported on the device -> "all-gather" -> device merge
        cands = []
        for s in hs:
            lib.check(L.tiktak_cuda_choose(s._h, None, None), "choose")
            c = torch.empty((K, 2), dtype=torch.float64, device=dev)
            lib.check(L.tiktak_cuda_choose_candidates(s._h, c.data_ptr()), "cand")
            cands.append(c)
        torch.cuda.synchronize()
        allc = torch.cat(cands).contiguous()
        assert int(allc.view(world, K, 2)[:, K - 1, 0].sum().item()) == o.n_legit  # trailer rows carry the valid counts
        for s in hs:
            xs = np.empty((nx + 1, K)); idx = np.zeros(K, dtype=np.int32)
            lib.check(L.tiktak_cuda_choose_merge(s._h, allc.data_ptr(), allc.shape[0], xs.ctypes.data, idx.ctypes.data), "merge")
            assert np.array_equal(xs.T, oxs) and np.array_equal(idx, o.sobol_idx)
        # --- restart rounds
        algid = 1 if alg == "a" else 0
        cap = D.n_max(cfg.round_size, world) * (nx + 2)
        sends = [torch.zeros(cap, dtype=torch.float64, device=dev) for _ in hs]
        for first_k, n_k in hs[0].rounds():
            cnt = D.n_max(n_k, world) * (nx + 2)
            for s, snd in zip(hs, sends):
                lib.check(L.tiktak_cuda_local_enqueue(s._h, algid, first_k, n_k, snd.data_ptr()), "enqueue")
            torch.cuda.synchronize()
            recv = torch.cat([snd[:cnt] for snd in sends]).contiguous()
            # the numpy restatement of the layout agrees with what the device packed
            got = D.unpack_rows_ref(recv.cpu().numpy().reshape(world * D.n_max(n_k, world), nx + 2), n_k, world)
            assert np.array_equal(got[:, 0], o.searchResults[first_k - 1: first_k - 1 + n_k, 3])  # values >= 2^-14: f40.20 exact
            for s in hs:
                lib.check(L.tiktak_cuda_local_ingest(s._h, first_k, n_k, recv.data_ptr()), "ingest")
            torch.cuda.synchronize()
        for s in hs:
            st = np.empty((nx, K)); r = np.empty((nx + 4, K)); e = np.zeros(K, dtype=np.int32)
            lib.check(L.tiktak_cuda_local_fetch(s._h, 1, K, st.ctypes.data, r.ctypes.data, e.ctypes.data), "fetch")
            assert np.array_equal(st.T, o.searchStart)
            assert np.array_equal(r.T, o.searchResults)
            assert np.array_equal(e, o.evals)
            assert s.getBestLine()[0] == o.getBestLine()[0]
    finally:
        for s in hs:
            s.close()


@pytest.mark.parametrize("name,alg,B,lp", [("griewank10", "a", 1, 3), ("griewank10", "a", 6, 10), ("rastrigin20", "a", 4, 2),
                                          ("griewank7", "b", 3, 4)])
def test_lottery_start_selection(built, name, alg, B, lp):
    """selectType = 1: getLotteryPoint (:722-766) with the counter-based random numbers of include/tiktak_rng.h"""
    kw = dict(round_size=B, search_type=1, lottery_points=lp)
    if alg == "b":
        kw.update(nm=3, maxpoints=9)
    cfg = cfg_of(name, **kw)
    with TikTakSearch(cfg) as s:
        s.solveAtSobolPoints(); s.chooseSobol(); s.LocalMinimizations(alg)
        gs, gr, ge = s.searchStart, s.searchResults, s.evals
    o = Oracle(cfg, MATH_TK)
    o.solveAtSobolPoints(); o.chooseSobol(mode=1); o.LocalMinimizations(1 if alg == "a" else 0)
    assert np.array_equal(gs, o.searchStart), "lottery-blended start points differ"
    assert np.array_equal(gr, o.searchResults) and np.array_equal(ge, o.evals)
    o0 = Oracle(cfg_of(name, **dict(kw, search_type=0)), MATH_TK)
    o0.solveAtSobolPoints(); o0.chooseSobol(mode=1); o0.LocalMinimizations(1 if alg == "a" else 0)
    assert not np.array_equal(o0.searchStart, o.searchStart)  # and it is not the best-point rule


def _edit_config(src, dst, repl):
    """rewrite scalar lines of a config.txt fixture: keys are the 1-based scalar positions (genericParams.f90:274-305)"""
    out, pos = [], 0
    for line in open(src):
        if not line.startswith("!") and pos < 9:
            pos += 1
            if pos in repl:
                line = f"{repl[pos]}   ! edited\n"
        out.append(line)
    open(dst, "w").writelines(out)


def test_cli_update_options_2_and_3(built, tmp_path):
    """driver options 2 (more Sobol points) and 3 (more local searches) resume from the .dat files of an earlier run
    (genericParams.f90:118-251)"""
    import os, shutil, subprocess
    from tiktak_b200 import parse_config
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "driver", "TiktakGlobalSearch")
    base = os.path.join(root, "tests", "golden", "config_c1.txt")
    run = lambda cwd, *a: subprocess.run([exe, *a], cwd=cwd, capture_output=True, text=True, timeout=300)
    a, b = tmp_path / "a", tmp_path / "b"
    a.mkdir(); b.mkdir()
    # ---- option 2: 20 -> 48 draws, legitimate 10 -> 30.  Same files as a cold start with the new config.
    cfg2 = str(tmp_path / "cfg2.txt")
    _edit_config(base, cfg2, {4: 48, 5: 30})
    assert run(a, "0", base, "a").returncode == 0
    r = run(a, "2", cfg2, "a")
    assert r.returncode == 0, r.stdout + r.stderr
    assert run(b, "0", cfg2, "a").returncode == 0
    for fn in ("sobolFnVal.dat", "x_starts.dat", "searchResults.dat", "searchStart.dat", "FinalResults.dat", "lastSobol.dat", "legitSobol.dat"):
        assert open(a / fn).read() == open(b / fn).read(), fn
    # an update that asks for fewer points than are solved is refused like the reference does (:189-195)
    r = run(a, "2", base, "a")
    assert r.returncode != 0 and int(open(a / "state.dat").read()) == -1
    # ---- option 3: 3 -> 6 local searches on top of run b's files
    cfg3 = str(tmp_path / "cfg3.txt")
    _edit_config(base, cfg3, {4: 48, 5: 30, 7: 6})
    before = open(b / "searchResults.dat").read().splitlines()
    r = run(b, "3", cfg3, "a")
    assert r.returncode == 0, r.stdout + r.stderr
    after = open(b / "searchResults.dat").read().splitlines()
    assert after[: len(before)] == before and len(after) == 1 + 7  # k = 0 record + restarts 1..7 (maxpoints + 1)
    rows = np.loadtxt(b / "searchResults.dat")
    cfg = parse_config(cfg3, objective_id=OBJ_TEMPLATE, round_size=1)
    o = Oracle(cfg, MATH_TK)
    o.solveAtSobolPoints(); o.chooseSobol(mode=1)
    o.LocalMinimizationsResume(1, rows[1:5], round_size=1)
    assert np.allclose(np.loadtxt(b / "x_starts.dat"), o.x_starts, rtol=0, atol=1e-15)
    assert np.array_equal(rows[1:, 1:3], o.searchResults[:, 1:3])  # k and the #improvements column continue
    assert np.allclose(rows[1:], o.searchResults, rtol=1e-15, atol=1e-19)
    starts = np.loadtxt(b / "searchStart.dat")
    assert np.allclose(starts[5:, 2:], o.searchStart[4:], rtol=1e-15, atol=1e-19)
    of, ox, _ = o.lastSearch()
    fin = np.loadtxt(b / "FinalResults.dat")
    assert abs(fin[2] - of) <= 1e-6 * abs(of)
    # option 3 with a different number of Sobol points is refused (:205-211)
    r = run(b, "3", base, "a")
    assert r.returncode != 0 and int(open(b / "state.dat").read()) == -1


@pytest.mark.parametrize("name,alg,B", [("griewank10", "a", 4), ("rastrigin20", "a", 3), ("griewank7", "b", 2)])
def test_early_started_rounds_equal_sequential_rounds(built, name, alg, B):
    """rounds started early against the current best row (tiktak_cuda_local_ingest_rounds) are accepted only while that
    row stays the best: any launch size gives the rows of strictly sequential rounds, and needs fewer launches"""
    kw = dict(round_size=B)
    if alg == "b":
        kw.update(nm=3, maxpoints=9)
    cfg = cfg_of(name, **kw)
    runs = {}
    for window in (0, 2 * B, 5 * B + 1, None):
        with TikTakSearch(cfg) as s:
            s.solveAtSobolPoints(); s.chooseSobol(); s.LocalMinimizations(alg, window=window)
            runs[window] = (s.searchStart.copy(), s.searchResults.copy(), s.evals.copy(), s.local_launches, s.getBestLine()[0])
    o = Oracle(cfg, MATH_TK)
    o.solveAtSobolPoints(); o.chooseSobol(mode=1); o.LocalMinimizations(1 if alg == "a" else 0)
    for window, (st, rs, ev, nl, fb) in runs.items():
        assert np.array_equal(st, o.searchStart) and np.array_equal(rs, o.searchResults) and np.array_equal(ev, o.evals), window
        assert fb == o.getBestLine()[0]
    n_rounds = -(-cfg.p_maxpoints // B)
    assert runs[0][3] == n_rounds and runs[None][3] < n_rounds


def _run_both(cfg, alg="a"):
    with TikTakSearch(cfg) as s:
        _, ne, nl = s.solveAtSobolPoints()
        fv = s.sobolFnVal() if ne > 0 else np.empty(0)
        xs = s.chooseSobol()
        idx = s.sobol_idx.copy()
        s.LocalMinimizations(alg)
        out = (ne, nl, fv, xs, idx, s.searchStart, s.searchResults, s.evals, s.getBestLine())
    o = Oracle(cfg, MATH_TK)
    _, one, onl = o.solveAtSobolPoints()
    oxs = o.chooseSobol(mode=1)
    o.LocalMinimizations(1 if alg == "a" else 0)
    return out, o, (one, onl, oxs)


def test_edge_no_sobol_points(built):
    """qr_ndraw = 0 -> maxpoints = 0 (genericParams.f90:325-326): one local search from p_init, nothing else"""
    cfg = make_config(nx=4, nm=2, qr_ndraw=0, maxpoints=5, lo=-1.0, hi=1.0, bound=[[-3.0, 3.0]] * 4, init=[0.8, -0.2, 0.3, 0.2],
                      objective_id=OBJ_TEMPLATE)
    assert cfg.p_maxpoints == 1
    (ne, nl, fv, xs, idx, st, rs, ev, best), o, (one, onl, oxs) = _run_both(cfg)
    assert (ne, nl) == (one, onl) == (0, 0)
    assert np.array_equal(xs, oxs) and xs.shape == (1, 5)
    assert np.array_equal(st, o.searchStart) and np.array_equal(rs, o.searchResults) and np.array_equal(ev, o.evals)


@pytest.mark.parametrize("alg", ["a", "b"])
def test_edge_one_parameter(built, alg):
    """nx = 1: a two-vertex simplex / npt = 3 interpolation points"""
    cfg = make_config(nx=1, nm=2, qr_ndraw=300, maxpoints=6, lo=-40.0, hi=60.0, init=[10.0], objective_id=OBJ_GRIEWANK, round_size=3)
    (ne, nl, fv, xs, idx, st, rs, ev, best), o, (one, onl, oxs) = _run_both(cfg, alg)
    assert (ne, nl) == (one, onl) and np.array_equal(fv, o.sobolFnVal())
    assert np.array_equal(xs, oxs) and np.array_equal(idx, o.sobol_idx)
    assert np.array_equal(st, o.searchStart) and np.array_equal(rs, o.searchResults) and np.array_equal(ev, o.evals)


def test_edge_largest_supported_dimension(built):
    """nx = 128 (the build's limit; runtime-nx Sobol kernel, fallback Nelder-Mead kernel with a short evaluation cap)"""
    cfg = make_config(nx=128, nm=1, qr_ndraw=600, maxpoints=3, lo=-2.0, hi=2.0, init=[0.5] * 128, objective_id=OBJ_ROSENBROCK,
                      round_size=2, nm_itmax=400)
    (ne, nl, fv, xs, idx, st, rs, ev, best), o, (one, onl, oxs) = _run_both(cfg)
    assert (ne, nl) == (one, onl) and np.array_equal(fv, o.sobolFnVal())
    assert np.array_equal(xs, oxs) and np.array_equal(idx, o.sobol_idx)
    assert np.array_equal(st, o.searchStart) and np.array_equal(rs, o.searchResults) and np.array_equal(ev, o.evals)


def test_edge_every_point_selected_and_none_legitimate(built):
    """maxpoints = qr_ndraw (every evaluated point becomes a start) and fvalmax below every value (no legitimate point):
    selection ignores validity (:882), invalid rows simply take part"""
    cfg = cfg_of("griewank7", qr_ndraw=37, maxpoints=37, fvalmax=1.0e-3, legitimate=5, round_size=8)
    (ne, nl, fv, xs, idx, st, rs, ev, best), o, (one, onl, oxs) = _run_both(cfg)
    assert (ne, nl) == (one, onl) == (37, 0)
    assert sorted(idx[1:].tolist()) == list(range(1, 38))
    assert np.array_equal(xs, oxs) and np.array_equal(idx, o.sobol_idx)
    assert np.array_equal(st, o.searchStart) and np.array_equal(rs, o.searchResults) and np.array_equal(ev, o.evals)


@pytest.mark.parametrize("alg", ["a", "b"])
def test_missing_search_sweep(built, alg):
    """States 8-9 (setupMissingSearch / SolveMissingSearch, TiktakGlobalSearch.f90:1048-1146): restarts that a killed instance
    never reported are run afterwards, one by one, each from getModifiedParam(p_maxpoints, x_starts(k,2:)) -- the blend weight
    of the last restart number toward the best row of the file as it is then -- and appended with #improvements counted
    against every row known.  Rows of a complete run minus three restarts are pushed back; the sweep must reproduce what the
    oracle's SolveMissingSearch does with the same file."""
    cfg = cfg_of("griewank7", round_size=4, maxpoints=19) if alg == "a" else cfg_of("griewank7", round_size=4, maxpoints=9, nm=3)
    K = cfg.p_maxpoints
    missing = [3, 7, K]
    algid = 1 if alg == "a" else 0
    with TikTakSearch(cfg) as s:
        s.solveAtSobolPoints(); xs = s.chooseSobol(); s.LocalMinimizations(alg)
        full = s.searchResults.copy()
    keep = full[[k - 1 for k in range(1, K + 1) if k not in missing]]
    with TikTakSearch(cfg) as s:
        s.solveAtSobolPoints(); s.chooseSobol()
        s.pushResults(keep)
        got = s.SolveMissingSearch(alg, missing)
        fb, xb, kb = s.getBestLine()
    o = Oracle(cfg, MATH_TK)
    o.solveAtSobolPoints(); o.chooseSobol(mode=1)
    row0 = np.concatenate([[1.0, 0.0, 0.0], xs[0]])  # the k = 0 record: [seqNo, 0, 0, fval(p_init), p_init]
    o.set_rows(np.vstack([row0, keep]))
    for k, (st, row, ev) in zip(missing, got):
        ost, orow, oev = o.SolveMissingSearch(algid, k)
        assert np.array_equal(st, ost) and np.array_equal(row, orow) and ev == oev, (k, row[:4], orow[:4], ev, oev)
    assert fb == o.getBestLine()[0]


def test_nelder_mead_on_a_vector_valued_objective(built):
    """amoeba through objFun = DOT_PRODUCT(dfovec, dfovec) + getPenalty (objective.f90:125-142) where dfovec fills a genuine
    vector of moments (the synthetic SMM: 600 moments from a simulated panel): one restart per block, block-wide evaluations"""
    cfg = smm_config(np_persons=1024, T=20, qr_ndraw=200, legitimate=30, maxpoints=5, fvalmax=20000.0, round_size=3)
    cfg.nm_itmax = 60  # an evaluation simulates the panel: keep the oracle's share of the test short
    with TikTakSearch(cfg) as s:
        s.solveAtSobolPoints(); s.chooseSobol(); s.LocalMinimizations("a")
        st, rs, ev, best = s.searchStart, s.searchResults, s.evals, s.getBestLine()
    o = Oracle(cfg, MATH_TK)
    o.solveAtSobolPoints(); o.chooseSobol(mode=1); o.LocalMinimizations(1)
    assert np.array_equal(st, o.searchStart) and np.array_equal(rs, o.searchResults) and np.array_equal(ev, o.evals)
    assert best[0] == o.getBestLine()[0] and ev.min() >= cfg.nx + 2


def test_local_minimizations_dfpmin(built):
    """the CLI's third algorithm, `d`: completeSearch CASE (2) = EST_dfpmin from the blended start (TiktakGlobalSearch.f90:575-576,
    minimize.f90:2572-2816), every objective evaluation on the device -- rows, starts and evaluation counts vs the oracle"""
    cfg = cfg_of("griewank7", round_size=3, maxpoints=8)
    with TikTakSearch(cfg) as s:
        s.solveAtSobolPoints(); s.chooseSobol(); s.LocalMinimizations("d")
        st, rs, ev, best = s.searchStart, s.searchResults, s.evals, s.getBestLine()
    o = Oracle(cfg, MATH_TK)
    o.solveAtSobolPoints(); o.chooseSobol(mode=1); o.LocalMinimizations(2)
    assert np.array_equal(st, o.searchStart) and np.array_equal(rs, o.searchResults) and np.array_equal(ev, o.evals)
    assert best[0] == o.getBestLine()[0]


def test_nan_objective_values_sort_last(built):
    """An objective that returns NaN on part of the box (here: tk_cos outside its domain |x| < 2^30 -- sign bit set or not)
    must not win the start selection: every NaN has one radix key above +inf, so numbers come first in (fval, index) order
    and NaN rows are taken, by ascending index, only when fewer than K-1 numbers exist."""
    for maxpoints in (6, 400):
        cfg = make_config(2, 1, 3000, maxpoints, -4.0e9, 4.0e9, init=[1.0, 2.0], objective_id=OBJ_GRIEWANK)
        with TikTakSearch(cfg) as s:
            _, ne, nl = s.solveAtSobolPoints()
            fv = s.sobolFnVal()
            xs = s.chooseSobol()
            idx = s.sobol_idx[1:].astype(np.int64)
        nan = np.isnan(fv)
        assert nan.any() and (~nan).sum() > 6 and nl == int((fv < cfg.fvalmax).sum())
        n = np.arange(1, ne + 1)
        order = np.concatenate([n[~nan][np.lexsort((n[~nan], fv[~nan]))], n[nan]])[:maxpoints]
        assert np.array_equal(idx, order), (maxpoints, idx[:10], order[:10])
        k_num = min(maxpoints, int((~nan).sum()))
        assert np.array_equal(xs[1:1 + k_num, 0], fv[order[:k_num] - 1]) and np.isnan(xs[1 + k_num:, 0]).all()


def test_edge_invalid_configurations_are_refused(built):
    """the reference stops (`stop 0`) on these; the library returns a status instead"""
    from tiktak_b200.lib import TikTakError
    cfg = cfg_of("griewank7")
    cfg.maxpoints = cfg.qr_ndraw + 1  # maxpoints > # sobol points (genericParams.f90:331-333)
    with pytest.raises(TikTakError):
        TikTakSearch(cfg)
    cfg = cfg_of("griewank7", search_type=2)  # unknown selectionMethod (TiktakGlobalSearch.f90:709-715)
    with pytest.raises(TikTakError):
        TikTakSearch(cfg)
    cfg = baseline_config("C3", sobol_quirk=1)  # 1e8 points: the literal reference sequence degenerates after 33,554,435 -- not built
    with pytest.raises(TikTakError):
        TikTakSearch(cfg)
    with TikTakSearch(cfg_of("griewank7", sobol_quirk=1)) as s:  # below that index both sequences coincide: accepted
        assert np.array_equal(s.setupSobol(1, 64), Oracle(cfg_of("griewank7", sobol_quirk=1), MATH_TK).setupSobol(1, 64, quirk=1))
    cfg = cfg_of("griewank7")
    with TikTakSearch(cfg) as s:
        with pytest.raises(TikTakError):
            s.chooseSobol()  # stage out of order
        s.solveAtSobolPoints()
        with pytest.raises(TikTakError):
            s.LocalMinimizations("a")  # before chooseSobol


def test_full_size_c2_against_oracle(built):
    """BASELINE.json configs[1] at full size (Griewank 10-D, 1e6 Sobol points, 1001 Nelder-Mead restarts in rounds of
    128): every stage bit for bit against the CPU oracle"""
    cfg = baseline_config("C2")
    cfg.round_size = 128
    with TikTakSearch(cfg) as s:
        _, ne, nl = s.solveAtSobolPoints()
        f = s.sobolFnVal()
        xs = s.chooseSobol()
        idx = s.sobol_idx.copy()
        s.LocalMinimizations("a")
        st, rs, ev, best = s.searchStart, s.searchResults, s.evals, s.getBestLine()
        ff, xf, it = s.lastSearch()
    o = Oracle(cfg, MATH_TK)
    _, one, onl = o.solveAtSobolPoints()
    assert (ne, nl) == (one, onl) == (1_000_000, 1_000_000)
    assert np.array_equal(f, o.sobolFnVal())
    assert np.array_equal(xs, o.chooseSobol(mode=1)) and np.array_equal(idx, o.sobol_idx)
    o.LocalMinimizations(1)
    assert np.array_equal(st, o.searchStart) and np.array_equal(rs, o.searchResults) and np.array_equal(ev, o.evals)
    ob = o.getBestLine()
    assert best[0] == ob[0] and best[2] == ob[2]
    of, ox, oit = o.lastSearch()
    assert abs(ff - of) <= 1e-6 * abs(of) and ff == of and np.array_equal(xf, ox)


@pytest.mark.parametrize("name,npts", [("C4", 10_000_000), ("C3", 100_000_000)])
def test_full_size_selection_properties(built, name, npts):
    """BASELINE.json configs[2] and [3] at full size, where the serial oracle would take minutes: size-independent
    properties of the Sobol stage and of the start selection.
      - the selected rows are in ascending (fval, index) order and no unselected point beats the last one;
      - every selected row is the Sobol point of its index (regenerated) and its fval is objFun of that point
        (re-evaluated through a different kernel): generation, evaluation and selection agree with one another;
      - the valid count equals the number of stored values below fvalmax;
      - selection is idempotent."""
    cfg = baseline_config(name)
    assert cfg.qr_ndraw == npts
    K, nx = cfg.p_maxpoints, cfg.nx
    with TikTakSearch(cfg) as s:
        _, ne, nl = s.solveAtSobolPoints()
        assert ne == npts
        xs = s.chooseSobol()
        idx = s.sobol_idx.copy()
        xs2 = s.chooseSobol()
        assert np.array_equal(xs, xs2) and np.array_equal(idx, s.sobol_idx)
        fsel, sel = xs[1:, 0], idx[1:].astype(np.int64)
        assert len(set(sel.tolist())) == K - 1 and sel.min() >= 1 and sel.max() <= npts
        order = np.lexsort((sel, fsel))
        assert np.array_equal(order, np.arange(K - 1))  # ascending fval, ties by ascending Sobol index
        # threshold: count of stored values <= the last selected value, chunk by chunk
        last_f, below, legit = fsel[-1], 0, 0
        chunk = 10_000_000
        for first in range(1, npts + 1, chunk):
            f = s.sobolFnVal(first, min(chunk, npts - first + 1))
            below += int(np.count_nonzero(f < last_f))
            legit += int(np.count_nonzero(f < cfg.fvalmax))
            inside = sel[(sel >= first) & (sel < first + len(f))]
            assert np.array_equal(f[inside - first], fsel[np.isin(sel, inside)])  # the stored value of every selected index
        assert below <= K - 2 and legit == nl
        # regenerate the selected points one by one and re-evaluate them
        for r in range(0, K - 1, max(1, (K - 1) // 200)):
            pt = s.setupSobol(int(sel[r]), 1)[0]
            assert np.array_equal(pt, xs[1 + r, 1:])
        assert np.array_equal(s.objFun(xs[1:, 1:]), fsel)


def test_full_size_local_stage_properties_c4(built):
    """BASELINE.json configs[3] at full size (Rosenbrock 50-D, 20 001 Nelder-Mead restarts): properties of the restart
    stage that hold for any size -- the reported value is objFun at the reported point, evaluation counts respect
    amoeba's cap, blended starts stay inside the Sobol box, the #improvements column is the running count of strict
    improvements over the k = 0 record, and getBestLine is the first minimum."""
    cfg = baseline_config("C4")
    cfg.round_size = 1184
    K, nx = cfg.p_maxpoints, cfg.nx
    with TikTakSearch(cfg) as s:
        s.solveAtSobolPoints(wait=False)
        xs = s.chooseSobol()
        s.LocalMinimizations("a")
        st, rs, ev = s.searchStart, s.searchResults, s.evals
        fb, xb, kb = s.getBestLine()
        f_again = s.objFun(rs[:, 4:])
    assert np.array_equal(rs[:, 1], np.arange(1, K + 1))
    assert np.array_equal(f_again, rs[:, 3])  # f40.20 keeps values >= 2^-14 exactly
    assert ev.min() >= nx + 2 and ev.max() <= cfg.nm_itmax + 2 * nx + 3
    lo, hi = np.asarray(cfg.range)[:, 0], np.asarray(cfg.range)[:, 1]
    assert np.all(st >= lo - 1e-12) and np.all(st <= hi + 1e-12)
    run_best, imp, want = xs[0, 0], 0, []
    for f in rs[:, 3]:
        if f < run_best:
            imp, run_best = imp + 1, f
        want.append(imp)
    assert np.array_equal(rs[:, 2], np.asarray(want, dtype=float))
    k_first_min = int(np.argmin(rs[:, 3])) + 1
    assert (fb, kb) == ((rs[k_first_min - 1, 3], k_first_min) if rs[k_first_min - 1, 3] < xs[0, 0] else (xs[0, 0], 0))
    assert np.array_equal(xb, rs[kb - 1, 4:] if kb > 0 else xs[0, 1:])


def test_smm_full_moment_count_properties(built):
    """BASELINE.json configs[4] shape (1200 moments, 7 parameters, T = 40; panel of 32768 persons): the `legitimate` cut
    stops at the first index where the valid count is reached, stored values are objFun of the regenerated points, and
    the DFNLS minima are re-evaluated to the reported values"""
    cfg = smm_config(np_persons=32768, T=40, qr_ndraw=2000, legitimate=20, maxpoints=15, fvalmax=50.0, round_size=16)
    assert cfg.nm == 1200
    with TikTakSearch(cfg) as s:
        _, ne, nl = s.solveAtSobolPoints()
        f = s.sobolFnVal()
        assert nl == 20 and ne < 2000 and len(f) == ne
        valid = f < cfg.fvalmax
        assert int(valid.sum()) == 20 and valid[-1] and int(valid[:-1].sum()) == 19  # the prefix ends at the 20th valid point
        pts = s.setupSobol(1, ne)
        probe = np.linspace(0, ne - 1, 40).astype(int)
        assert np.array_equal(s.objFun(pts[probe]), f[probe])
        xs = s.chooseSobol()
        assert np.array_equal(np.sort(xs[1:, 0]), xs[1:, 0]) and xs[1, 0] == f.min()
        s.LocalMinimizations("b")
        rs, ev = s.searchResults, s.evals
        assert np.array_equal(s.objFun(rs[:, 4:]), rs[:, 3])
        assert ev.max() <= cfg.maxeval + 1 and ev.min() >= 2 * cfg.nx + 2
        assert rs[:, 3].min() <= xs[:, 0].min()  # the local stage does not lose the best start
