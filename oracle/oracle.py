"""ctypes wrapper of the CPU oracle (oracle/libtiktak_oracle.so).  TEST INFRASTRUCTURE ONLY.

May be imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs -- never by the tiktak_b200 package.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from tiktak_b200._abi import TiktakConfig, as_f64, c_double_p, c_int32_p, dptr, smm_user

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libtiktak_oracle.so")
_lib = None


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE])


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        L = C.CDLL(LIB_PATH)
        L.orc_create.restype = C.c_void_p
        L.orc_create.argtypes = [C.POINTER(TiktakConfig), C.c_int]
        L.orc_destroy.argtypes = [C.c_void_p]
        L.orc_penalty.restype = C.c_double
        L.orc_penalty.argtypes = [C.c_void_p, c_double_p]
        L.orc_text4020.restype = C.c_double
        L.orc_text4020.argtypes = [C.c_double]
        L.orc_w11.restype = C.c_float
        L.orc_nfev.restype = C.c_long
        L.orc_nfev.argtypes = [C.c_void_p]
        L.orc_best_tie_events.restype = C.c_long
        L.orc_best_tie_events.argtypes = [C.c_void_p]
        for name in ("orc_sobol_points", "orc_objfun", "orc_dfovec", "orc_pretest", "orc_pretest_values", "orc_choose",
                     "orc_set_starts", "orc_local", "orc_best", "orc_run_amoeba", "orc_dfnls"):
            getattr(L, name).restype = C.c_int
        L.orc_sobol_points.argtypes = [C.c_void_p, C.c_int64, C.c_int64, c_double_p, C.c_int, C.c_int]
        L.orc_sobol_unit.argtypes = [C.c_int, C.c_int, C.c_int64, c_double_p, C.c_int]
        L.orc_sobol_dirnums.argtypes = [C.c_int, C.c_int, C.c_void_p]
        L.orc_objfun.argtypes = [C.c_void_p, c_double_p, C.c_int64, c_double_p]
        L.orc_dfovec.argtypes = [C.c_void_p, c_double_p, c_double_p]
        L.orc_pretest.argtypes = [C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
        L.orc_pretest_values.argtypes = [C.c_void_p, C.c_int64, C.c_int64, c_double_p]
        L.orc_choose.argtypes = [C.c_void_p, C.c_int, c_double_p, c_int32_p, C.POINTER(C.c_int)]
        L.orc_set_starts.argtypes = [C.c_void_p, c_double_p]
        L.orc_local.argtypes = [C.c_void_p, C.c_int, C.c_int, c_double_p, c_double_p, c_int32_p]
        L.orc_best.argtypes = [C.c_void_p, c_double_p, c_double_p, c_int32_p]
        L.orc_run_amoeba.argtypes = [C.c_void_p, c_double_p, c_int32_p]
        L.orc_dfnls.argtypes = [C.c_void_p, c_double_p, C.c_double, C.c_double, C.c_int, c_int32_p, c_int32_p]
        L.orc_set_dfnls_force_rescue.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.orc_set_dfnls_force_rescue.restype = C.c_long
        L.orc_dfnls_rescue_evals.argtypes = [C.c_void_p]
        L.orc_dfnls_rescue_evals.restype = C.c_long
        L.orc_simplex_coordinates2.argtypes = [C.c_int, c_double_p]
        L.orc_indexx.argtypes = [C.c_int, c_double_p, c_int32_p]
        L.orc_philox4x32_10.argtypes = [C.POINTER(C.c_uint32), C.c_uint32, C.c_uint32]
        L.orc_philox4x32_10.restype = None
        L.orc_uniform01.argtypes = [C.c_uint32, C.c_uint32, C.c_uint64]
        L.orc_uniform01.restype = C.c_double
        _lib = L
    return _lib


MATH_LIBM, MATH_TK = 0, 1


class Oracle:
    """Single-instance CPU run of the reference algorithm (see oracle/tiktak_oracle.cpp)."""

    def __init__(self, cfg, math_mode=MATH_TK):
        self.cfg = cfg
        self.L = load()
        self._range = as_f64(np.asarray(cfg.range).T.reshape(-1))
        self._init = as_f64(cfg.init)
        self._bound = as_f64(np.asarray(cfg.bound).T.reshape(-1))
        self._user = smm_user(cfg)
        self._c = TiktakConfig(
            nx=cfg.nx, nm=cfg.nm, maxeval=cfg.maxeval, qr_ndraw=cfg.qr_ndraw, legitimate=cfg.legitimate,
            maxpoints_plus1=cfg.p_maxpoints, search_type=cfg.search_type, lottery_points=cfg.lottery_points,
            fvalmax=cfg.fvalmax, range=dptr(self._range), init=dptr(self._init), bound=dptr(self._bound),
            objective_id=cfg.objective_id, device=0, rank=0, world=1, nm_itmax=cfg.nm_itmax,
            nm_shrink_mode=cfg.nm_shrink_mode, sobol_quirk=cfg.sobol_quirk, text_roundtrip=cfg.text_roundtrip,
            round_size=cfg.round_size, nm_ftol=cfg.nm_ftol, user=C.cast(C.pointer(self._user), C.c_void_p) if self._user is not None else None)
        self.h = C.c_void_p(self.L.orc_create(C.byref(self._c), math_mode))
        assert self.h, "orc_create failed"

    def close(self):
        if self.h:
            self.L.orc_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def K(self):
        return self.cfg.p_maxpoints

    def setupSobol(self, first=1, count=None, quirk=0, atmost=0):
        count = self.cfg.qr_ndraw if count is None else count
        out = np.empty((self.cfg.nx, count))
        rc = self.L.orc_sobol_points(self.h, first, count, dptr(out), quirk, atmost or self.cfg.qr_ndraw or 1)
        assert rc == 0, rc
        return out.T

    def objFun(self, theta):
        theta = np.atleast_2d(np.asarray(theta, dtype=np.float64))
        x = np.ascontiguousarray(theta.T)
        f = np.empty(theta.shape[0])
        assert self.L.orc_objfun(self.h, dptr(x), theta.shape[0], dptr(f)) == 0
        return f

    def solveAtSobolPoints(self):
        ne, nl = C.c_int64(), C.c_int64()
        rc = self.L.orc_pretest(self.h, C.byref(ne), C.byref(nl))
        assert rc == 0, rc
        self.n_evaluated, self.n_legit = ne.value, nl.value
        return True, ne.value, nl.value

    def sobolFnVal(self, first=1, count=None):
        count = self.n_evaluated if count is None else count
        f = np.empty(count)
        self.L.orc_pretest_values(self.h, first, count, dptr(f))
        return f

    def chooseSobol(self, mode=0):
        """mode 0 = indexx (reference), 1 = stable (fval, index) order."""
        K, nx = self.K, self.cfg.nx
        xs = np.empty((nx + 1, K))
        idx = np.zeros(K, dtype=np.int32)
        tie = C.c_int()
        rows = self.L.orc_choose(self.h, mode, dptr(xs), idx.ctypes.data_as(c_int32_p), C.byref(tie))
        assert rows >= 0
        self.x_starts, self.sobol_idx, self.tie_straddle, self.n_start_rows = xs.T.copy(), idx, bool(tie.value), rows
        return self.x_starts

    def set_starts(self, x_starts):
        xs = np.ascontiguousarray(np.asarray(x_starts, dtype=np.float64).T)
        assert self.L.orc_set_starts(self.h, dptr(xs)) == 0
        self.x_starts = np.asarray(x_starts, dtype=np.float64).copy()

    def LocalMinimizations(self, algor=1, round_size=None):
        K, nx = self.K, self.cfg.nx
        B = self.cfg.round_size if round_size is None else round_size
        s = np.empty((nx, K)); r = np.empty((nx + 4, K)); e = np.zeros(K, dtype=np.int32)
        rc = self.L.orc_local(self.h, algor, B, dptr(s), dptr(r), e.ctypes.data_as(c_int32_p))
        assert rc == 0, rc
        self.searchStart, self.searchResults, self.evals = s.T.copy(), r.T.copy(), e
        return True

    def LocalMinimizationsAsync(self, instances, algor=1, clock=0):
        """the asynchronous-instances run (Oracle::local_async); sets searchStart / searchResults / evals / fin / inst.
        clock (Nelder-Mead): 0 = a tick per pass of amoeba's loop, 1 = a tick per objective evaluation (tiktak_cuda_local_async_clock)"""
        K, nx = self.K, self.cfg.nx
        self.L.orc_set_async_clock.argtypes = [C.c_void_p, C.c_int]
        self.L.orc_set_async_clock.restype = None
        self.L.orc_set_async_clock(self.h, int(clock))
        s = np.empty((nx, K)); r = np.empty((nx + 4, K)); e = np.zeros(K, dtype=np.int32)
        fin = np.zeros(K, dtype=np.int32); inst = np.zeros(K, dtype=np.int32)
        self.L.orc_local_async.argtypes = [C.c_void_p, C.c_int, C.c_int, c_double_p, c_double_p, c_int32_p, c_int32_p, c_int32_p]
        rc = self.L.orc_local_async(self.h, algor, instances, dptr(s), dptr(r), e.ctypes.data_as(c_int32_p), fin.ctypes.data_as(c_int32_p),
                                    inst.ctypes.data_as(c_int32_p))
        assert rc == 0, rc
        self.searchStart, self.searchResults, self.evals, self.fin, self.inst = s.T.copy(), r.T.copy(), e, fin, inst
        return True

    def LocalMinimizationsResume(self, algor, rows, round_size=None):
        """option 3: `rows` (n, nx+4) = the finished restarts 1..n as re-read from searchResults.dat; runs n+1..K"""
        K, nx = self.K, self.cfg.nx
        B = self.cfg.round_size if round_size is None else round_size
        rin = np.ascontiguousarray(np.asarray(rows, dtype=np.float64).T)
        s = np.full((nx, K), np.nan); r = np.empty((nx + 4, K)); e = np.zeros(K, dtype=np.int32)
        self.L.orc_local_resume.argtypes = [C.c_void_p, C.c_int, C.c_int, c_double_p, C.c_int, c_double_p, c_double_p, c_int32_p]
        rc = self.L.orc_local_resume(self.h, algor, B, dptr(rin), len(rows), dptr(s), dptr(r), e.ctypes.data_as(c_int32_p))
        assert rc == 0, rc
        self.searchStart, self.searchResults, self.evals = s.T.copy(), r.T.copy(), e
        return True

    def set_rows(self, rows):
        """rows (n, nx+4) = the records of an interrupted run in file order, the k = 0 record first"""
        rin = np.ascontiguousarray(np.asarray(rows, dtype=np.float64).T)
        self.L.orc_set_rows.argtypes = [C.c_void_p, c_double_p, C.c_int]
        assert self.L.orc_set_rows(self.h, dptr(rin), len(rows)) == 0

    def SolveMissingSearch(self, algor, k):
        """states 8-9 for one missing restart k (TiktakGlobalSearch.f90:1104-1146): (start, searchResults row, evaluations)"""
        nx = self.cfg.nx
        st, row, ne = np.empty(nx), np.empty(nx + 4), C.c_int32()
        self.L.orc_missing_search.argtypes = [C.c_void_p, C.c_int, C.c_int, c_double_p, c_double_p, c_int32_p]
        assert self.L.orc_missing_search(self.h, algor, k, dptr(st), dptr(row), C.byref(ne)) == 0
        return st, row, ne.value

    def getBestLine(self):
        f, k = C.c_double(), C.c_int32()
        x = np.empty(self.cfg.nx)
        assert self.L.orc_best(self.h, C.byref(f), dptr(x), C.byref(k)) == 0
        return f.value, x, k.value

    def lastSearch(self):
        f, it = C.c_double(), C.c_int32()
        x = np.empty(self.cfg.nx)
        self.L.orc_last_search.argtypes = [C.c_void_p, c_double_p, c_double_p, c_int32_p]
        assert self.L.orc_last_search(self.h, C.byref(f), dptr(x), C.byref(it)) == 0
        return f.value, x, it.value

    def run_amoeba(self, start):
        p = as_f64(start).copy()
        it = C.c_int32()
        self.L.orc_run_amoeba(self.h, dptr(p), C.byref(it))
        return p, it.value

    def sobolFnValRange(self, first, count):
        """objFun at Sobol points first..first+count-1 (seeks by the Gray code; no points before `first` are generated)"""
        f = np.empty(count)
        self.L.orc_objfun_range.restype = C.c_long
        self.L.orc_objfun_range.argtypes = [C.c_void_p, C.c_int64, C.c_int64, c_double_p]
        nl = self.L.orc_objfun_range(self.h, first, count, dptr(f))
        assert nl >= 0, nl
        return f, int(nl)

    def sobolPointsRange(self, first, count):
        out = np.empty((self.cfg.nx, count))
        self.L.orc_points_range.argtypes = [C.c_void_p, C.c_int64, C.c_int64, c_double_p]
        assert self.L.orc_points_range(self.h, first, count, dptr(out)) == 0
        return out.T

    def completeSearch(self, algor, k, start):
        """one restart from a given start: (result point, fn_val of the re-evaluation :582, objective evaluations)"""
        x = as_f64(start).copy()
        f = C.c_double()
        self.L.orc_complete_search.restype = C.c_int
        self.L.orc_complete_search.argtypes = [C.c_void_p, C.c_int, C.c_int, c_double_p, c_double_p]
        ne = self.L.orc_complete_search(self.h, algor, k, dptr(x), C.byref(f))
        assert ne >= 0, ne
        return x, f.value, ne

    def nfev(self):
        return int(self.L.orc_nfev(self.h))
