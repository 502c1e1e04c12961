"""CPU tests: the oracle's local searches pinned by INDEPENDENT pure-Python restatements of the Fortran
(tests/golden/ref_restate.py: amoeba/amotry minimize.f90:7-135, runAmoeba TiktakGlobalSearch.f90:602-633,
EST_dfpmin/dograd/lnsrch minimize.f90:2572-2816, objFun/dfovec/getPenalty objective.f90:51-142).

The restatements share no text with oracle/tiktak_oracle.cpp or the CUDA kernels, run on Python floats (IEEE double,
no FMA) and call the same glibc `cos` as the oracle's MATH_LIBM mode, so every comparison below is bit for bit:
minima, vertices, evaluation counts.  A transcription slip in the oracle (and therefore in anything validated
against it) shows up here.
"""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
import ref_restate as R  # noqa: E402

from oracle.oracle import MATH_LIBM, Oracle  # noqa: E402
from tiktak_b200 import (OBJ_GRIEWANK, OBJ_RASTRIGIN, OBJ_ROSENBROCK, OBJ_TEMPLATE, baseline_config,  # noqa: E402
                         make_config)

KIND = {OBJ_TEMPLATE: "template", OBJ_GRIEWANK: "griewank", OBJ_RASTRIGIN: "rastrigin", OBJ_ROSENBROCK: "rosenbrock"}


def restated_objective(cfg):
    return R.Objective(KIND[cfg.objective_id], cfg.nm, cfg.bound[:, 0], cfg.bound[:, 1], cfg.fvalmax)


CASES = [
    ("C1 template", lambda: baseline_config("C1")),
    ("griewank10", lambda: make_config(10, 1, 3000, 12, -400.0, 600.0, init=[100.0] * 10, objective_id=OBJ_GRIEWANK)),
    ("rastrigin6 nm3", lambda: make_config(6, 3, 2000, 8, -4.12, 5.12, init=[1.0] * 6, objective_id=OBJ_RASTRIGIN)),
    ("rosenbrock5", lambda: make_config(5, 1, 2000, 8, -5.0, 10.0, init=[2.5] * 5, objective_id=OBJ_ROSENBROCK)),
    # a box much smaller than the simplex: vertices are clamped, trial points leave the bounds -> penalty branch
    ("griewank4 tight bounds", lambda: make_config(4, 2, 500, 6, -50.0, 70.0, init=[5.0] * 4, bound=[[-6.0, 9.0]] * 4,
                                                   fvalmax=80.0, objective_id=OBJ_GRIEWANK)),
]


@pytest.mark.parametrize("name,mk", CASES, ids=[c[0] for c in CASES])
def test_objfun_bit_exact(built, name, mk):
    cfg = mk()
    o = Oracle(cfg, MATH_LIBM)
    f = restated_objective(cfg)
    rng = np.random.default_rng(5)
    lo, hi = cfg.bound[:, 0], cfg.bound[:, 1]
    pts = rng.uniform(lo - 0.3 * (hi - lo), hi + 0.3 * (hi - lo), size=(400, cfg.nx))  # a third of them outside the box
    got = o.objFun(pts)
    for x, g in zip(pts, got):
        assert f([float(v) for v in x]) == g


@pytest.mark.parametrize("name,mk", CASES, ids=[c[0] for c in CASES])
def test_amoeba_bit_exact(built, name, mk):
    cfg = mk()
    o = Oracle(cfg, MATH_LIBM)
    f = restated_objective(cfg)
    rng = np.random.default_rng(11)
    lo, hi = cfg.range[:, 0], cfg.range[:, 1]
    starts = [cfg.init] + [rng.uniform(lo, hi) for _ in range(5)]
    n_oob = 0
    for st in starts:
        p_o, it_o = o.run_amoeba(st)
        nf0 = f.nfev
        p_r, it_r = R.run_amoeba([float(v) for v in st], f, lo.tolist(), hi.tolist(), cfg.bound[:, 0].tolist(),
                                 cfg.bound[:, 1].tolist(), cfg.p_maxpoints, cfg.nm_ftol, cfg.nm_itmax)
        assert it_o == it_r, (name, it_o, it_r)
        assert np.array_equal(p_o, np.asarray(p_r)), name
        assert f.nfev - nf0 == cfg.nx + 1 + it_r  # nx+1 vertices + `iter` evaluations
        n_oob += int(np.any(np.asarray(p_r) <= cfg.bound[:, 0]) or np.any(np.asarray(p_r) >= cfg.bound[:, 1]))
    if "tight" in name:
        assert n_oob > 0  # the clamped / penalised path was really taken


def test_amoeba_every_branch_is_exercised(built):
    """the restated amoeba and the oracle agree on cases that take every branch, the shrink (minimize.f90:106-113) included"""
    cfg = make_config(3, 1, 200, 4, -4.12, 5.12, init=[0.3, -0.2, 0.1], objective_id=OBJ_RASTRIGIN)
    o = Oracle(cfg, MATH_LIBM)
    f = restated_objective(cfg)
    stats = {}
    rng = np.random.default_rng(3)
    for _ in range(40):
        st = rng.uniform(-4.0, 5.0, 3)
        p_r, it_r = R.run_amoeba(st.tolist(), f, cfg.range[:, 0].tolist(), cfg.range[:, 1].tolist(), cfg.bound[:, 0].tolist(),
                                 cfg.bound[:, 1].tolist(), cfg.p_maxpoints, stats=stats)
        p_o, it_o = o.run_amoeba(st)
        assert it_o == it_r and np.array_equal(p_o, np.asarray(p_r))
    assert all(stats.get(k, 0) > 0 for k in ("reflect", "expand", "contract", "shrink")), stats


@pytest.mark.parametrize("name,mk", CASES[:4], ids=[c[0] for c in CASES[:4]])
def test_dfpmin_bit_exact(built, name, mk):
    """lastSearch (TiktakGlobalSearch.f90:635-681): EST_dfpmin from the best searchResults row"""
    cfg = mk()
    cfg.round_size = 1
    o = Oracle(cfg, MATH_LIBM)
    o.solveAtSobolPoints(); o.chooseSobol(); o.LocalMinimizations(1)
    fb, xb, kb = o.getBestLine()
    f_o, x_o, it_o = o.lastSearch()
    f = restated_objective(cfg)
    x_r, f_r, it_r = R.est_dfpmin([float(v) for v in xb], f)
    assert it_o == it_r, (it_o, it_r)
    assert f_o == f_r and np.array_equal(x_o, np.asarray(x_r))
    assert f_r <= fb


# ------------------------------------------------------------------------------------------- DFNLS (BOBYQA_H family)
# The oracle's DFNLS (oracle/dfnls_oracle.inc, 1,360 lines) and the CUDA kernel follow the same Fortran statement by statement,
# so comparing one with the other cannot show a slip made in both.  The pins below share no text with either:
#   * known answers: the minimiser of least-squares problems with a known / independently computed (scipy, numpy) solution;
#   * invariants of Powell's method that every correct iteration must keep, checked with numpy at EVERY trust-region step of
#     whole runs: the per-moment quadratic models interpolate the residuals at all NPT points (GQV/HQV/PQV updates, XBASE
#     shifts, PRELIM_H), and BMAT/ZMAT hold the inverse KKT matrix, i.e. the Lagrange functions they define are 1 at their own
#     point and 0 at the others (UPDATE, the shift formulas of label 90, PRELIM_H's initial factorisation).
import ctypes as C  # noqa: E402

from oracle.oracle import MATH_TK  # noqa: E402
from tiktak_b200._abi import c_double_p, c_int32_p, dptr  # noqa: E402

OBJ_TEST_ROSENRES, OBJ_TEST_LINRES = 100, 101  # oracle-only residual objectives (oracle/tiktak_oracle.cpp dfovec)


def _dfnls_cfg(obj, nx, nm, lo, hi, init):
    return make_config(nx, nm, 10, 1, lo, hi, init=init, objective_id=obj, fvalmax=1.0e8)


def _run_dfnls(cfg, start, rhobeg, rhoend, maxfun, trace=0, force=None, info=None):
    o = Oracle(cfg, MATH_TK)
    if force:
        o.L.orc_set_dfnls_force_rescue(o.h, force[0], force[1])
    x = np.ascontiguousarray(np.asarray(start, dtype=np.float64))
    nf = C.c_int32()
    n, npt, mv = cfg.nx, 2 * cfg.nx + 1, cfg.nm
    if not trace:
        nr = C.c_int32()
        rc = o.L.orc_dfnls(o.h, dptr(x), rhobeg, rhoend, maxfun, C.byref(nf), C.byref(nr))
        assert rc == 0
        if info is not None:
            info.update(nrescue=nr.value, rescue_evals=o.L.orc_dfnls_rescue_evals(o.h))
        return x, nf.value, None
    nh, ndim, nptm = n * (n + 1) // 2, npt + n, npt - n - 1
    size = 2 + 2 * n + npt * n + mv * npt + mv * n + mv * nh + mv * npt + ndim * n + npt * nptm
    buf = np.zeros(trace * size)
    o.L.orc_dfnls_trace.restype = C.c_int
    o.L.orc_dfnls_trace.argtypes = [C.c_void_p, c_double_p, C.c_double, C.c_double, C.c_int, C.c_int, c_double_p, c_int32_p]
    ns = o.L.orc_dfnls_trace(o.h, dptr(x), rhobeg, rhoend, maxfun, trace, dptr(buf), C.byref(nf))
    if info is not None:
        info.update(nrescue=o.L.orc_set_dfnls_force_rescue(o.h, 0, 0), rescue_evals=o.L.orc_dfnls_rescue_evals(o.h))
    snaps = []
    for s in range(ns):
        w = buf[s * size:(s + 1) * size]
        pos = [2]

        def take(rows, cols):
            a = w[pos[0]:pos[0] + rows * cols].reshape(cols, rows).T.copy()  # Fortran column-major (rows, cols)
            pos[0] += rows * cols
            return a
        snaps.append(dict(kopt=int(w[0]), nf=int(w[1]), xbase=take(n, 1)[:, 0], xopt=take(n, 1)[:, 0], xpt=take(npt, n), fval_v=take(mv, npt),
                          gqv=take(mv, n), hqv=take(mv, nh), pqv=take(mv, npt), bmat=take(ndim, n), zmat=take(npt, nptm)))
    return x, nf.value, snaps


def _rosen_res(x, extra=0):
    x = np.asarray(x)
    r = np.full(2 * (len(x) - 1) + extra, 0.5)
    r[0:2 * (len(x) - 1):2] = 10.0 * (x[1:] - x[:-1] ** 2)
    r[1:2 * (len(x) - 1):2] = 1.0 - x[:-1]
    return r


def _lin_Ab(nx, nm):
    m, j = np.arange(1, nm + 1)[:, None], np.arange(1, nx + 1)[None, :]
    return np.sin(1.3 * m + 0.7 * j + 0.1 * m * j), 0.5 + 0.25 * np.cos(1.0 + np.arange(nm))


def test_dfnls_known_answers(built):
    from scipy.optimize import least_squares

    # Rosenbrock as residuals (the classic test of least-squares codes) plus one constant residual 0.5: the minimiser is
    # (1, ..., 1) with F = 0.25
    for nx, start in ((2, [-1.2, 1.0]), (4, [-1.2, 1.0, -1.2, 1.0])):
        cfg = _dfnls_cfg(OBJ_TEST_ROSENRES, nx, 2 * (nx - 1) + 1, -5.0, 5.0, start)
        x, nf, _ = _run_dfnls(cfg, start, 0.5, 1.0e-9, 5000)
        ref = least_squares(lambda v: _rosen_res(v, 1), start, bounds=(-5.0, 5.0), xtol=1e-14, ftol=1e-14, gtol=1e-14).x
        assert np.allclose(x, 1.0, atol=1e-6) and np.allclose(x, ref, atol=1e-6), (nx, x, ref, nf)
    # the zero-residual form runs into the reference's own early exit `F <= max(1e-12, 1e-20 FBEG)` (minimize.f90:831-835), which
    # jumps to label 720 BEFORE the new point is stored: the search returns the best EARLIER point, one trust-region step away
    # from the minimiser it has just found.  A quirk of the reference (kept): the returned point is not (1, 1), its successor was.
    cfg = _dfnls_cfg(OBJ_TEST_ROSENRES, 2, 2, -5.0, 5.0, [-1.2, 1.0])
    x, nf, _ = _run_dfnls(cfg, [-1.2, 1.0], 0.5, 1.0e-9, 5000)
    assert nf < 100 and 1e-12 < float((_rosen_res(x) ** 2).sum()) < 0.1 and np.allclose(x, 1.0, atol=0.05)
    # linear least squares: unconstrained solution inside the box (numpy lstsq) ...
    nx, nm = 3, 8
    A, b = _lin_Ab(nx, nm)
    xs = np.linalg.lstsq(A, b, rcond=None)[0]
    cfg = _dfnls_cfg(OBJ_TEST_LINRES, nx, nm, -10.0, 10.0, [0.0] * nx)
    x, nf, _ = _run_dfnls(cfg, [0.0] * nx, 1.0, 1.0e-9, 2000)
    assert np.all(np.abs(xs) < 10.0) and np.allclose(x, xs, atol=1e-6), (x, xs, nf)
    # ... and with an active bound (scipy's bounded trust-region-reflective solver)
    hi = float(xs.max()) - 0.2
    cfg = _dfnls_cfg(OBJ_TEST_LINRES, nx, nm, -10.0, hi, [0.0] * nx)
    x, nf, _ = _run_dfnls(cfg, [min(0.0, hi - 0.5)] * nx, 0.25, 1.0e-9, 2000)
    ref = least_squares(lambda v: A @ v - b, [min(0.0, hi - 0.5)] * nx, bounds=(-10.0, hi), xtol=1e-14, ftol=1e-14, gtol=1e-14).x
    assert np.isclose(x.max(), hi) and np.allclose(x, ref, atol=1e-5), (x, ref, nf)


def _model_and_kkt_errors(cfg, snaps, resid):
    """largest violation, over the snapshots of a traced run, of (1) the interpolation conditions of the per-moment models and
    (2) the Lagrange property of the functions BMAT / ZMAT define"""
    n, npt = cfg.nx, 2 * cfg.nx + 1
    iu = [(i, j) for j in range(n) for i in range(j + 1)]  # packed upper triangle by columns (IH runs j = 1..n, i = 1..j)
    worst_i = worst_l = 0.0
    for s in snaps:
        xpt, ko = s["xpt"], s["kopt"] - 1
        assert np.array_equal(s["xopt"], xpt[ko])
        # the stored residual vectors are the objective's residuals at XBASE + XPT(k, :)
        for k in range(npt):
            assert np.allclose(s["fval_v"][:, k], resid(s["xbase"] + xpt[k]), rtol=0, atol=1e-9 * (1 + np.abs(s["fval_v"][:, k]).max()))
        # (1) interpolation: v_m(x_k) - v_m(x_opt) = g_m . d + d' H_m d / 2,  H_m = HQV_m + sum_j PQV_mj x_j x_j'
        d = xpt - xpt[ko]
        for m in range(cfg.nm):
            H = np.zeros((n, n))
            for ih, (i, j) in enumerate(iu):
                H[i, j] = H[j, i] = s["hqv"][m, ih]
            H = H + (xpt.T * s["pqv"][m]) @ xpt
            model = d @ s["gqv"][m] + 0.5 * np.einsum("ki,ij,kj->k", d, H, d)
            truth = s["fval_v"][m] - s["fval_v"][m, ko]
            scale = 1.0 + np.abs(s["fval_v"][m]).max()
            worst_i = max(worst_i, float(np.abs(model - truth).max() / scale))
        # (2) BMAT/ZMAT: Lagrange function k = c + BMAT(k,:).x + sum_i Omega_ki (x_i.x)^2 / 2 with Omega = ZMAT ZMAT' is the
        #     Kronecker delta on the interpolation points (differences taken against x_opt, so the constant drops out)
        omega = s["zmat"] @ s["zmat"].T
        q = 0.5 * (xpt @ xpt.T) ** 2          # q[i, j] = (x_i . x_j)^2 / 2
        lag = omega @ (q - q[:, [ko]]) + s["bmat"][:npt] @ (xpt - xpt[ko]).T   # [k, j] = l_k(x_j) - l_k(x_opt)
        want = np.eye(npt) - np.eye(npt)[:, [ko]]
        worst_l = max(worst_l, float(np.abs(lag - want).max()))
    return worst_i, worst_l


@pytest.mark.parametrize("case", ["rosen2", "rosen3", "lin3"])
def test_dfnls_models_interpolate_and_kkt_inverse_holds_at_every_step(built, case):
    if case.startswith("rosen"):
        nx = int(case[-1])
        cfg = _dfnls_cfg(OBJ_TEST_ROSENRES, nx, 2 * (nx - 1), -3.0, 3.0, [-1.2, 1.0, -0.5][:nx])
        start, resid, rb = [-1.2, 1.0, -0.5][:nx], _rosen_res, 0.4
    else:
        nx, nm = 3, 8
        A, b = _lin_Ab(nx, nm)
        cfg = _dfnls_cfg(OBJ_TEST_LINRES, nx, nm, -10.0, 10.0, [0.3, -0.2, 0.1])
        start, resid, rb = [0.3, -0.2, 0.1], (lambda v: A @ v - b), 0.5
    x, nf, snaps = _run_dfnls(cfg, start, rb, 1.0e-7, 600, trace=400)
    n, npt = cfg.nx, 2 * cfg.nx + 1
    assert len(snaps) >= 10 and nf > npt
    worst_i, worst_l = _model_and_kkt_errors(cfg, snaps, resid)
    assert worst_i < 1e-7, worst_i
    assert worst_l < 1e-6, worst_l


@pytest.mark.parametrize("case,first,every", [("rosen3", 9, 0), ("rosen3", 12, 5), ("lin3", 10, 6), ("lin3", 8, 0),
                                              ("rosen3", 8, 0), ("rosen5", 15, 0), ("rosen5", 24, 0)])
def test_dfnls_rescue_keeps_the_interpolation_system(built, case, first, every):
    """RESCUE_H (minimize.f90:1627-2023) has no counterpart in scipy and almost never runs on its own, so the oracle's
    restatement is pinned through what it must leave behind.  With the test hook forcing label 190 at chosen evaluation
    numbers: after EVERY later step BMAT / ZMAT -- which RESCUE_H rebuilds from scratch, exchanging provisional for original
    points one UPDATE at a time -- still define Lagrange functions on the current points, whether or not provisional points
    stayed (the last three cases: evaluations inside RESCUE_H, its second half :1892-2019), and the per-moment models still
    interpolate the residuals at every point.  The run ends where the plain run ends."""
    if case.startswith("rosen"):
        nx = int(case[-1])
        start = [-1.2, 1.0, -0.5, 0.8, -1.0][:nx]
        cfg = _dfnls_cfg(OBJ_TEST_ROSENRES, nx, 2 * (nx - 1) + 1, -3.0, 3.0, start)
        resid, rb = (lambda v: _rosen_res(v, 1)), 0.4
    else:
        nx, nm = 3, 8
        A, b = _lin_Ab(nx, nm)
        cfg = _dfnls_cfg(OBJ_TEST_LINRES, nx, nm, -10.0, 10.0, [0.3, -0.2, 0.1])
        start, resid, rb = [0.3, -0.2, 0.1], (lambda v: A @ v - b), 0.5
    info = {}
    x, nf, snaps = _run_dfnls(cfg, start, rb, 1.0e-7, 900, trace=600, force=(first, every), info=info)
    assert info["nrescue"] >= 1 and len(snaps) >= 10
    assert (info["rescue_evals"] > 0) == ((case, first) in (("rosen3", 8), ("rosen5", 15), ("rosen5", 24)))
    worst_i, worst_l = _model_and_kkt_errors(cfg, snaps, resid)
    assert worst_l < 1e-6, worst_l
    assert worst_i < 1e-7, worst_i
    plain, _, _ = _run_dfnls(cfg, start, rb, 1.0e-7, 900)
    assert np.allclose(x, plain, atol=2e-4), (x, plain)
