"""CPU tests: the oracle (oracle/tiktak_oracle.cpp) against known answers.

The reference ships no tests or golden files, so the pins are
 (1) the known-answer vectors derived from the reference's formulas, and
 (2) tests/golden/golden_small.json, produced by an independent pure-Python restatement
     (tests/golden/make_golden.py) that shares no code with the C++ oracle.
"""
import ctypes as C
import json
import os

import numpy as np
import pytest

from oracle import oracle as orc
from oracle.oracle import MATH_LIBM, MATH_TK, Oracle
from tiktak_b200 import OBJ_GRIEWANK, OBJ_RASTRIGIN, OBJ_ROSENBROCK, baseline_config, make_config
from tiktak_b200._abi import c_double_p, c_int32_p, dptr

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "golden_small.json")))


@pytest.fixture(scope="module")
def L(built):
    return orc.load()


def unit_points(L, nx, n, atmost=None, quirk=0):
    out = np.zeros((n, nx))
    assert L.orc_sobol_unit(nx, atmost or max(n, 1), n, dptr(out), quirk) == 0
    return out


def test_no_fma_contraction(L):
    assert L.orc_unfused_selfcheck() == 0


def test_sobol_first_points_survey(L):
    exp = np.array([[.5, .5, .5, .5], [.75, .25, .75, .25], [.25, .75, .25, .75], [.375, .375, .625, .125],
                    [.875, .875, .125, .625], [.625, .125, .375, .375], [.125, .625, .875, .875],
                    [.1875, .3125, .3125, .6875]])
    assert np.array_equal(unit_points(L, 4, 8, 20), exp)


def test_sobol_matches_independent_python(L):
    assert np.array_equal(unit_points(L, 4, 16), np.array(GOLD["sobol_unit_nx4_first16"]))
    assert np.array_equal(unit_points(L, 10, 200), np.array(GOLD["sobol_unit_nx10_first200"]))
    p50 = unit_points(L, 50, 40)
    assert np.array_equal(p50[[0, 1, 7, 38, 39]], np.array(GOLD["sobol_unit_nx50_sample"]))


def test_sobol_independent_of_atmost(L):
    a = unit_points(L, 10, 500, atmost=500)
    b = unit_points(L, 10, 500, atmost=(1 << 30) - 1)
    assert np.array_equal(a, b)
    assert L.orc_sobol_maxcol(4, 20) == 5  # C1: maxcol = 5


def test_sobol_dims_1_2_match_scipy(L):
    qmc = pytest.importorskip("scipy.stats.qmc")
    ref = qmc.Sobol(d=2, scramble=False, bits=30).random(257)[1:]  # scipy starts at the zero vector, Gray order
    # scipy >= 1.7 generates in Gray-code order as well; dims 1-2 share direction numbers with Bratley-Fox
    ours = unit_points(L, 2, 256)
    assert np.array_equal(np.sort(ours, axis=0), np.sort(ref, axis=0))
    assert np.array_equal(ours, ref)


def test_sobol_seek_by_gray_code_equals_sequential_generation(built):
    """orc_objfun_range / orc_points_range start in the middle of the sequence (state = XOR of the columns picked by the
    Gray code of the previous index): same points and values as generating from point 1 -- what lets the full-size
    GPU tests share 1e8 oracle evaluations between processes"""
    cfg = make_config(20, 1, 70000, 10, -4.12, 5.12, init=[1.0] * 20, objective_id=OBJ_RASTRIGIN)
    o = Oracle(cfg, MATH_TK)
    o.solveAtSobolPoints()
    f, pts = o.sobolFnVal(), o.setupSobol()
    for first, count in [(1, 50), (2, 3), (4096, 10), (4097, 200), (32768, 5), (65535, 3), (69990, 11), (33333, 1)]:
        g, nl = o.sobolFnValRange(first, count)
        assert np.array_equal(g, f[first - 1:first - 1 + count]) and nl == count
        assert np.array_equal(o.sobolPointsRange(first, count), pts[first - 1:first - 1 + count])


def test_sobol_single_precision_quirk(L):
    """utilities.f90:1090 `i = floor(real(i/2))`: identical to the integer rule up to Scount 33,554,434,
    first mismatch at 33,554,435 (l_ref = 2, l_true = 3), then only columns 1..2 are ever used."""
    for s in [0, 1, 2, 3, 7, 1023, 2**20 - 1, 2**24 + 1, 33554431, 33554433, 33554434]:
        assert L.orc_sobol_l(s, 1) == L.orc_sobol_l(s, 0), s
    assert (L.orc_sobol_l(33554435, 1), L.orc_sobol_l(33554435, 0)) == (2, 3)
    rng = np.random.default_rng(0)
    for s in rng.integers(33554435, 2**30 - 1, size=2000):
        lt = L.orc_sobol_l(int(s), 0)
        assert L.orc_sobol_l(int(s), 1) == min(lt, 2)


def test_c1_pretest_and_choose(built):
    cfg = baseline_config("C1")
    o = Oracle(cfg, MATH_LIBM)
    _, ne, nl = o.solveAtSobolPoints()
    assert (ne, nl) == (10, 10)  # only the first 10 of 20 points are evaluated (legitimate = 10)
    pts = o.setupSobol(1, 10)
    assert np.array_equal(pts, np.array(GOLD["c1_points"]))
    f = o.sobolFnVal()
    assert np.array_equal(f, np.array(GOLD["c1_fvals"]))  # bit-exact vs the independent restatement
    surv = [8.0, 12.306743316801496, 12.306743316801496, 10.133087507051615, 16.595497939018895, 11.19727818240388,
            15.688444943098277, 12.562292117323677, 12.655206982365566, 18.96798569747271]
    assert np.allclose(f, surv, rtol=3e-16, atol=0)  # SURVEY's figures used 8*v^2, the code sums v^2 eight times
    assert f[1] == f[2]  # mirror points tie exactly
    xs = o.chooseSobol(mode=0)
    assert o.sobol_idx.tolist() == [0, 1, 4, 6] and not o.tie_straddle
    assert xs[0, 0] == GOLD["c1_finit"] and abs(xs[0, 0] - 14.258672740927095) < 1e-14
    assert o.objFun([3.5, 0, 0, 0])[0] == GOLD["c1_penalty_case"] == 4145.6790123456785


def test_tk_math_vs_libm_objective(built):
    for name in ("C1",):
        cfg = baseline_config(name)
        a, b = Oracle(cfg, MATH_LIBM), Oracle(cfg, MATH_TK)
        a.solveAtSobolPoints(); b.solveAtSobolPoints()
        assert np.max(np.abs(a.sobolFnVal() - b.sobolFnVal()) / a.sobolFnVal()) <= 1e-12
    for oid, nx, lo, hi in ((OBJ_GRIEWANK, 10, -400, 600), (OBJ_RASTRIGIN, 20, -4.12, 5.12)):
        cfg = make_config(nx, 1, 20000, 10, lo, hi, objective_id=oid)
        a, b = Oracle(cfg, MATH_LIBM), Oracle(cfg, MATH_TK)
        a.solveAtSobolPoints(); b.solveAtSobolPoints()
        assert np.max(np.abs(a.sobolFnVal() - b.sobolFnVal()) / a.sobolFnVal()) <= 1e-12


def test_blend_weights_single_precision(L):
    assert [float(L.orc_w11(k, 4)) for k in range(1, 5)] == GOLD["w11_K4"] == [
        0.5, 0.7071067690849304, 0.8660253882408142, 0.949999988079071]
    assert [float(L.orc_w11(k, 1001)) for k in (1, 2, 10, 11, 500, 903, 904, 1001)] == GOLD["w11_K1001"]


def test_simplex_coordinates2(L):
    for n, key in ((4, "simplex4"), (10, "simplex10")):
        x = np.zeros((n + 1, n))
        L.orc_simplex_coordinates2(n, dptr(x))
        g = np.array(GOLD[key])  # (n, n+1)
        assert np.allclose(x.T, g, rtol=0, atol=2e-16)
        assert np.allclose(np.linalg.norm(x, axis=1), 1.0, atol=1e-15)
    x = np.zeros((5, 4)); L.orc_simplex_coordinates2(4, dptr(x))
    assert abs(x[0, 0] - 0.9635254915624212) < 1e-15 and abs(x[0, 1] + 0.1545084971874737) < 1e-15
    assert abs(x[4, 0] + 0.5000000000000001) < 1e-15


def test_indexx_is_a_sort_and_insertion_part_is_stable(L):
    rng = np.random.default_rng(1)
    for n in (1, 2, 14, 15, 16, 17, 100, 5000):
        a = rng.standard_normal(n)
        idx = np.zeros(n, dtype=np.int32)
        assert L.orc_indexx(n, dptr(a), idx.ctypes.data_as(c_int32_p)) == 0
        assert np.array_equal(idx - 1, np.argsort(a, kind="stable"))  # no ties: any correct sort agrees
    a = np.array([3.0, 1.0, 3.0, 1.0, 2.0, 3.0])  # n < 15 -> pure insertion sort, stable (utilities.f90:1355 `<=`)
    idx = np.zeros(6, dtype=np.int32)
    L.orc_indexx(6, dptr(a), idx.ctypes.data_as(c_int32_p))
    assert idx.tolist() == [2, 4, 5, 1, 3, 6]


def test_text_roundtrip_matches_arithmetic_version(L):
    """snprintf/strtod ('200f40.20' records) vs include/tiktak_math.h's closed form used on the GPU"""
    src = r'''
#include "tiktak_math.h"
double f(double x) { return tk_f4020_roundtrip(x); }
double c(double x) { return tk_cos(x); }
double s(double x) { return tk_sin(x); }
'''
    import subprocess, tempfile
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "t.c"), "w").write(src)
        so = os.path.join(d, "t.so")
        subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-mfma", "-shared", "-fPIC", "-I", os.path.join(root, "include"),
                               os.path.join(d, "t.c"), "-o", so, "-lm"])
        T = C.CDLL(so)
        for fn in (T.f, T.c, T.s):
            fn.restype = C.c_double; fn.argtypes = [C.c_double]
        rng = np.random.default_rng(2)
        xs = np.concatenate([rng.uniform(-1, 1, 4000) * 2.0 ** -rng.integers(0, 45, 4000),
                             [(2 * k + 1) / 2.0 ** 21 for k in range(64)], [0.0, -0.0, 1e-30, 123.456, -7e-5, 6.103515625e-05]])
        for x in xs:
            assert T.f(float(x)) == L.orc_text4020(float(x)) or (x == 0 and T.f(float(x)) == 0)
        # tk_cos / tk_sin against libm: <= 2 ulp on the ranges the objectives use
        xs = np.concatenate([rng.uniform(-700, 700, 20000), rng.uniform(-40, 40, 20000), rng.uniform(-1e5, 1e5, 5000)])
        for x in xs[:8000]:
            for mine, ref in ((T.c(float(x)), np.cos(x)), (T.s(float(x)), np.sin(x))):
                assert abs(mine - ref) <= 2.0 * np.spacing(abs(ref)) + 0.0


def test_nelder_mead_reaches_known_minima(built):
    """analytic minima: objFun = nm * (f + 1)^2 with f_min = 0 -> nm"""
    cfg = make_config(10, 1, 4000, 20, -400.0, 600.0, init=[100.0] * 10, objective_id=OBJ_GRIEWANK, round_size=1)
    o = Oracle(cfg, MATH_TK)
    o.solveAtSobolPoints(); o.chooseSobol(); o.LocalMinimizations(1)
    assert o.evals.max() <= 5000 + 10 + 2
    fb, xb, kb = o.getBestLine()
    assert 1.0 <= fb < 1.5
    cfg = make_config(4, 1, 2000, 10, -2.0, 2.0, init=[0.5] * 4, objective_id=OBJ_ROSENBROCK, round_size=1)
    o = Oracle(cfg, MATH_TK)
    o.solveAtSobolPoints(); o.chooseSobol(); o.LocalMinimizations(1)
    fb, xb, _ = o.getBestLine()
    assert abs(fb - 1.0) < 1e-3 and np.allclose(xb, 1.0, atol=0.05)


def test_round_schedule_and_counter_column(built):
    cfg = baseline_config("C1")
    o = Oracle(cfg, MATH_TK)
    o.solveAtSobolPoints(); o.chooseSobol()
    o.LocalMinimizations(1, round_size=1)
    seq = o.searchResults.copy(); st = o.searchStart.copy()
    # restart 1 starts at p_init itself; restart 2 is blended toward the best row so far
    assert np.array_equal(st[0], cfg.init)
    assert not np.array_equal(st[1], o.x_starts[1, 1:])
    assert list(seq[:, 1]) == [1, 2, 3, 4]
    assert np.all(np.diff(seq[:, 2]) >= 0)  # the #improvements counter never decreases
    o.LocalMinimizations(1, round_size=4)  # one round: nobody sees a finished search -> pure Sobol starts
    assert np.array_equal(o.searchStart[1:], o.x_starts[1:, 1:])


def test_asynchronous_instances_event_simulation(built):
    """Oracle::local_async, the model nm_quad_async_kernel / dfnls_kernel reproduce on the device, checked against its own
    definition with numpy: (1) one instance is the sequential run (rounds of one restart); (2) for P instances, replaying the
    schedule from the reported finish times and instance numbers gives back every start row: restart k starts when the
    earliest-free instance is free (lowest number among equals), sees exactly the rows with a finish time <= that, and is
    blended toward the lowest of them (ties in file order) with the single-precision weight of k; (3) the #improvements column
    counts along the file order (finish time, instance)."""
    cfg = make_config(nx=6, nm=2, qr_ndraw=900, maxpoints=29, lo=-40.0, hi=60.0, init=[10.0] * 6, objective_id=OBJ_GRIEWANK)
    K, nx = cfg.p_maxpoints, cfg.nx
    o1 = Oracle(cfg, MATH_TK)
    o1.solveAtSobolPoints(); o1.chooseSobol(mode=1); o1.LocalMinimizations(1, round_size=1)
    oa = Oracle(cfg, MATH_TK)
    oa.solveAtSobolPoints(); xs = oa.chooseSobol(mode=1); oa.LocalMinimizationsAsync(1)
    assert np.array_equal(o1.searchStart, oa.searchStart) and np.array_equal(o1.searchResults, oa.searchResults) and np.array_equal(o1.evals, oa.evals)
    assert np.all(np.diff(oa.fin) > 0) and np.all(oa.inst == 0)
    for alg, P, clock in ((1, 5, 0), (1, 12, 0), (0, 4, 0), (1, 5, 1), (1, 12, 1)):  # clock 1: the one-warp instances' (tick = evaluation)
        o = Oracle(cfg, MATH_TK)
        o.solveAtSobolPoints(); xs = o.chooseSobol(mode=1); o.LocalMinimizationsAsync(P, algor=alg, clock=clock)
        fin, inst, res, st = o.fin.astype(np.int64), o.inst, o.searchResults, o.searchStart
        f0 = xs[0, 0]                                    # the k = 0 record: the value of the initial guess
        free = np.zeros(P, dtype=np.int64)
        for k in range(1, K + 1):
            i = int(np.argmin(free))                     # earliest free instance, lowest number among equals
            assert inst[k - 1] == i
            T = free[i]
            vis = [j for j in range(1, k) if fin[j - 1] <= T]
            cand = [(f0, 0, -1, 0)] + [(res[j - 1, 3], fin[j - 1], inst[j - 1], j) for j in vis]
            best = min(cand)
            if len(cand) <= 1:
                want = xs[k - 1, 1:]
            else:
                w = np.float64(min(max(np.float32(0.10), np.sqrt(np.float32(k) / np.float32(K))), np.float32(0.95)))
                base = xs[0, 1:] if best[3] == 0 else res[best[3] - 1, 4:]
                want = w * base + (1.0 - w) * xs[k - 1, 1:]
            assert np.array_equal(st[k - 1], want), (alg, P, k)
            assert fin[k - 1] > T + 32                   # pickup ticks + at least one tick of work
            if clock == 1 or alg == 0:                   # a restart costs the pickup (DFNLS 32, one-warp Nelder-Mead 128 ticks) + its objective evaluations
                assert fin[k - 1] == T + (32 if alg == 0 else 128) + o.evals[k - 1]
            free[i] = fin[k - 1]
        order = np.lexsort((inst, fin))                  # file order: finish time, then instance
        best_f, imp = f0, 0
        for j in order:
            if res[j, 3] < best_f:
                best_f, imp = res[j, 3], imp + 1
            assert res[j, 2] == imp
        assert P == 1 or len(set(inst.tolist())) == P


def test_legitimate_cut_prefix_semantics(built):
    cfg = make_config(10, 1, 5000, 10, -400.0, 600.0, objective_id=OBJ_GRIEWANK)
    o = Oracle(cfg, MATH_TK); o.solveAtSobolPoints()
    f = o.sobolFnVal()
    med = float(np.median(f))
    cfg2 = make_config(10, 1, 5000, 10, -400.0, 600.0, objective_id=OBJ_GRIEWANK, fvalmax=med, legitimate=100)
    o2 = Oracle(cfg2, MATH_TK)
    _, ne, nl = o2.solveAtSobolPoints()
    # NB fvalmax also sets the penalty constants, but all Sobol points are inside the bounds
    valid = np.cumsum(f < med)
    assert nl == 100 and ne == int(np.argmax(valid >= 100)) + 1
    cfg3 = make_config(10, 1, 50, 10, -400.0, 600.0, objective_id=OBJ_GRIEWANK, fvalmax=1e-3, legitimate=5)
    o3 = Oracle(cfg3, MATH_TK)
    assert o3.solveAtSobolPoints() == (True, 50, 0)  # never reached: all points evaluated (soft goal)


def _philox_py(c, k0, k1):
    """independent restatement of Philox4x32-10 (Salmon et al., SC'11)"""
    c = list(c)
    M = 0xFFFFFFFF
    for _ in range(10):
        p0, p1 = 0xD2511F53 * c[0], 0xCD9E8D57 * c[2]
        c = [((p1 >> 32) ^ c[1] ^ k0) & M, p1 & M, ((p0 >> 32) ^ c[3] ^ k1) & M, p0 & M]
        k0, k1 = (k0 + 0x9E3779B9) & M, (k1 + 0xBB67AE85) & M
    return c


def test_lottery_random_numbers(L):
    """include/tiktak_rng.h: Philox4x32-10 known answers (Random123 kat_vectors) and the [0,1) mapping"""
    import ctypes as C
    kats = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
            ((0xffffffff,) * 4, (0xffffffff, 0xffffffff), (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
            ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0), (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, want in kats:
        c = (C.c_uint32 * 4)(*ctr)
        L.orc_philox4x32_10(c, key[0], key[1])
        assert tuple(c) == want == tuple(_philox_py(ctr, *key))
    rng = np.random.default_rng(5)
    for _ in range(50):
        seed, stream, ctr = int(rng.integers(0, 2**32)), int(rng.integers(0, 2**32)), int(rng.integers(0, 2**63))
        out = _philox_py((ctr & 0xFFFFFFFF, ctr >> 32, stream, 0), seed, 0)
        u = ((out[0] << 32 | out[1]) >> 11) * 2.0 ** -53
        assert L.orc_uniform01(seed, stream, ctr) == u and 0.0 <= u < 1.0


def test_lottery_selection_oracle(built):
    """selectType = 1 (getLotteryPoint :722-766): restarts blend toward a rank-weighted random earlier result, not only
    the best one; rank probabilities 6/11, 3/11, 2/11 at three points"""
    from tiktak_b200 import OBJ_GRIEWANK
    cfg = make_config(nx=10, nm=1, qr_ndraw=3000, maxpoints=40, lo=-400.0, hi=600.0, init=[100.0] * 10, objective_id=OBJ_GRIEWANK,
                      search_type=1, lottery_points=3, round_size=1)
    o = Oracle(cfg, MATH_TK); o.solveAtSobolPoints(); o.chooseSobol(mode=1); o.LocalMinimizations(1)
    cfg0 = make_config(nx=10, nm=1, qr_ndraw=3000, maxpoints=40, lo=-400.0, hi=600.0, init=[100.0] * 10, objective_id=OBJ_GRIEWANK,
                       round_size=1)
    o0 = Oracle(cfg0, MATH_TK); o0.solveAtSobolPoints(); o0.chooseSobol(mode=1); o0.LocalMinimizations(1)
    assert np.array_equal(o.searchStart[0], o0.searchStart[0])  # k = 1: only the k = 0 record exists, no blend (:701)
    assert not np.array_equal(o.searchStart, o0.searchStart)
    # reconstruct the blend of every restart from the rows known at its start: the base point must be one of the 3 best rows
    K, nx = cfg.p_maxpoints, cfg.nx
    rows = [(o.x_starts[0, 0], 0, o.x_starts[0, 1:])]
    hits = {1: 0, 2: 0, 3: 0}
    for k in range(1, K + 1):
        if len(rows) > 1:
            order = sorted(rows, key=lambda r: (r[0], r[1]))[:3]
            w = np.float64(min(max(np.float32(0.10), np.sqrt(np.float32(k) / np.float32(K))), np.float32(0.95)))
            cands = [w * r[2] + (1.0 - w) * o.x_starts[k - 1, 1:] for r in order]
            match = [i for i, c in enumerate(cands) if np.array_equal(c, o.searchStart[k - 1])]
            assert match, k
            hits[match[0] + 1] += 1
        rows.append((o.searchResults[k - 1, 3], k, o.searchResults[k - 1, 4:]))
    assert hits[1] > hits[3] and hits[2] > 0  # rank 1 is drawn most often, lower ranks do occur


def test_affine_map_identity_used_by_the_sobol_kernel():
    """kernels_sobol.cu maps a point with I2F, DMUL, DFMA: fma(fl(q * w), 2^-32, lo).  The reference computes
    lo + fl(u * w) with u = q * 2^-32 (TiktakGlobalSearch.f90:834).  Scaling by a power of two is exact and commutes
    with rounding (no underflow at these magnitudes), so both give the same double -- checked here bit for bit on the
    CPU for the BASELINE boxes and for awkward widths; the product q * w * 2^-32 being exact, the FMA equals a plain add."""
    rng = np.random.default_rng(20261020)
    q = rng.integers(1, 2**32, size=200_000, dtype=np.uint64)
    q[:64] = [1 << (i % 32) for i in range(64)]
    for lo, hi in ((-400.0, 600.0), (-4.12, 5.12), (-5.0, 10.0), (-1.0, 1.0), (1e-3, 1e-3 + 2.0**-20), (-3.0e9, 3.0e9),
                   (0.1, 0.3), (-1e-9, 1e-9)):
        w = np.float64(hi) - np.float64(lo)
        u = q.astype(np.float64) * np.float64(2.0**-32)            # exact
        ref = np.float64(lo) + u * w                                   # two roundings, as the reference
        dev = np.float64(lo) + (q.astype(np.float64) * w) * np.float64(2.0**-32)  # fl(q*w) scaled exactly, then one add
        assert np.array_equal(ref, dev), (lo, hi)


def test_sobol_kernel_table_algebra(L):
    """the index algebra of sobol_eval_kernel, restated with numpy on the oracle's direction numbers: a block of 128
    threads owning the indices [base, base + 128 P) seeds thread t with (XOR of the rows picked by the Gray bits >= 7 of
    base) ^ Lt[t], where Lt[t] = XOR of the rows picked by gray(t) (the same in every block because base is a multiple of
    256), and steps n -> n + 128 with Dt[c] = V[6] ^ V[7 + c], c = ctz(~(n >> 7)).  Every point produced that way must be
    the oracle's point (q * 2^-32 in the unit cube)."""
    nx, S, P, BT = 10, 7, 8, 128
    atmost = 2**30 - 1
    V = np.zeros((nx, 32), dtype=np.uint32)
    assert L.orc_sobol_dirnums(nx, atmost, V.ctypes.data) == 0
    V = V.astype(np.uint64)
    Lt = np.zeros((BT, nx), dtype=np.uint64)
    for t in range(BT):
        g = t ^ (t >> 1)
        for j in range(S):
            if (g >> j) & 1:
                Lt[t] ^= V[:, j]
    Dt = np.stack([V[:, S - 1] ^ V[:, S + c] for c in range(30 - S)])
    ref = unit_points(L, nx, 3 * BT * P + 5000, atmost=atmost)  # points 1 .. N
    for base in (0, BT * P, 2 * BT * P, 4096):
        h0 = base >> S
        ghi = h0 ^ (h0 >> 1)
        qhi = np.zeros(nx, dtype=np.uint64)
        for j in range(30 - S):
            if (ghi >> j) & 1:
                qhi ^= V[:, j + S]
        q = Lt ^ qhi[None, :]                                   # (thread, coordinate)
        hh = h0
        for j in range(P):
            n = base + j * BT + np.arange(BT)
            ok = n >= 1
            pts = q.astype(np.float64) * 2.0**-32
            assert np.array_equal(pts[ok], ref[n[ok] - 1]), (base, j)
            c = (~hh & (hh + 1)).bit_length() - 1               # ctz(~hh)
            q = q ^ Dt[c][None, :]
            hh += 1


def test_radix_select_with_list_restated():
    """the selection algorithm of kernels_select.cu restated with numpy: order-preserving 64-bit key, radix select most
    significant byte first, the list of elements sharing the threshold's top 16 bits after two passes, the remaining
    digits (and the tie-break by index) from the list only -- against a plain lexicographic sort by (value, index),
    for distributions with one bucket, many buckets, negative values and heavy ties"""
    def key(f):
        u = f.view(np.uint64)
        return np.where(u >> np.uint64(63) != 0, ~u, u | np.uint64(1 << 63))

    def select(f, m):
        k = key(f)
        n = np.arange(1, len(f) + 1, dtype=np.uint64)
        prefix, rem = np.uint64(0), m
        src_k, src_n = k, n
        for shift in range(56, -8, -8):
            sh = np.uint64(shift)
            match = np.ones(len(src_k), bool) if shift == 56 else (src_k >> (sh + np.uint64(8))) == (prefix >> (sh + np.uint64(8)))
            hist = np.bincount(((src_k[match] >> sh) & np.uint64(255)).astype(np.int64), minlength=256)
            cum = np.concatenate([[0], np.cumsum(hist)])
            b = int(np.searchsorted(cum[1:], rem, side="left"))
            b = min(b, 255)
            rem -= int(cum[b])
            prefix |= np.uint64(b) << sh
            if shift == 48:  # compaction: only these can still matter
                keep = (k >> np.uint64(48)) == (prefix >> np.uint64(48))
                src_k, src_n = k[keep], n[keep]
        eq = src_k == prefix
        n_thr = np.uint64(0xFFFFFFFF)
        if int(eq.sum()) > rem:  # more keys equal to the threshold than still needed: lowest indices win
            n_thr = np.sort(src_n[eq])[rem - 1]
        win = (k < prefix) | ((k == prefix) & (n <= n_thr))
        return np.flatnonzero(win)

    rng = np.random.default_rng(7)
    cases = [rng.random(50_000) * 1e-4 + 1.0, rng.lognormal(3.0, 4.0, 50_000), rng.normal(0.0, 10.0, 50_000),
             np.round(rng.random(50_000) * 20.0), np.ones(5_000)]
    for f in cases:
        f = f.astype(np.float64)
        for m in (1, 37, 1000, len(f) - 1, len(f)):
            got = select(f, m)
            order = np.lexsort((np.arange(len(f)), f))[:m]
            assert np.array_equal(np.sort(got), np.sort(order)), (len(f), m)
