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
