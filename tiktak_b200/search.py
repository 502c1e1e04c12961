"""Host-side mirror of the reference's stage procedures over the C ABI.

The method names, argument meaning and error behaviour follow the internal procedures of
PROGRAM TiktakGlobalSearch (TiktakGlobalSearch.f90:404-1148) for the hot path:

    setupSobol            :825-838     sobol_trial(i,:) = lo + q*(hi-lo)
    solveAtSobolPoints    :405-471     objFun at points 1.. until `legitimate` valid values
    chooseSobol           :840-899     x_starts = [p_init; best maxpoints Sobol points]
    LocalMinimizations    :473-544     restarts k = 1..p_maxpoints (getModifiedParam + completeSearch)
    getBestLine           :781-802     best searchResults row so far
    run                   :217-401     the state machine 1 -> 2 -> 3 -> 6 -> 7 (-> 10) of one cold start

All compute happens in libtiktak_cuda (CUDA, sm_100a).  torch is used only for device memory
hand-off and torch.distributed (NCCL) when world > 1: the Sobol index range and each round's
restarts are sharded over the ranks and the per-round result rows are all-gathered.
"""
import ctypes as C

import numpy as np

from . import lib as _lib
from ._abi import ALG_AMOEBA, ALG_DFNLS, SOBOL_BLOCK, TiktakConfig, as_f64, dptr, smm_user
from .config import TikTakConfig

_ALG = {"a": ALG_AMOEBA, "b": ALG_DFNLS, "amoeba": ALG_AMOEBA, "dfnls": ALG_DFNLS, 0: ALG_DFNLS, 1: ALG_AMOEBA}


class TikTakSearch:
    """One "instance" of the reference program bound to one GPU (rank of a torch.distributed group or alone)."""

    def __init__(self, cfg: TikTakConfig, device=0, rank=0, world=1, group=None):
        self.cfg = cfg
        self.rank, self.world, self.group = rank, world, group
        self.device = device
        self._L = _lib.load()
        # keep the arrays alive for the lifetime of the struct
        self._range = as_f64(np.asarray(cfg.range).T.reshape(-1))  # column-major (nx,2): lo.., hi..
        self._init = as_f64(cfg.init)
        self._bound = as_f64(np.asarray(cfg.bound).T.reshape(-1))
        self._user = smm_user(cfg)
        self._c = TiktakConfig(
            nx=cfg.nx, nm=cfg.nm, maxeval=cfg.maxeval, qr_ndraw=cfg.qr_ndraw, legitimate=cfg.legitimate,
            maxpoints_plus1=cfg.p_maxpoints, search_type=cfg.search_type, lottery_points=cfg.lottery_points,
            fvalmax=cfg.fvalmax, range=dptr(self._range), init=dptr(self._init), bound=dptr(self._bound),
            objective_id=cfg.objective_id, device=device, rank=rank, world=world, nm_itmax=cfg.nm_itmax,
            nm_shrink_mode=cfg.nm_shrink_mode, sobol_quirk=cfg.sobol_quirk, text_roundtrip=cfg.text_roundtrip,
            round_size=cfg.round_size, nm_ftol=cfg.nm_ftol, user=C.cast(C.pointer(self._user), C.c_void_p) if self._user is not None else None)
        self._h = C.c_void_p()
        _lib.check(self._L.tiktak_cuda_create(C.byref(self._h), C.byref(self._c)), "tiktak_cuda_create")
        self.n_evaluated = None
        self.n_legit = None
        self.x_starts = None
        self.sobol_idx = None
        self.searchStart = None  # rows of searchStart.dat: (K, nx)
        self.searchResults = None  # rows of searchResults.dat: (K, nx+4) = seqNo, k, #improv, fval, x
        self.evals = None

    # ------------------------------------------------------------------ lifetime
    def close(self):
        if self._h:
            self._L.tiktak_cuda_destroy(self._h)
            self._h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ helpers
    @property
    def K(self):
        return self.cfg.p_maxpoints

    def launch_count(self):
        return int(self._L.tiktak_cuda_launch_count(self._h))

    def stage_ms(self, stage):
        ms = C.c_double()
        _lib.check(self._L.tiktak_cuda_stage_ms(self._h, stage.encode(), C.byref(ms)), "stage_ms")
        return ms.value

    def _dist(self):
        import torch.distributed as dist

        return dist

    # ------------------------------------------------------------------ stage 2
    def setupSobol(self, first=1, count=None):
        """sobol_trial rows first..first+count-1, shape (count, nx) (the reference fills all qr_ndraw rows)."""
        count = self.cfg.qr_ndraw if count is None else count
        out = np.empty((self.cfg.nx, count), dtype=np.float64)  # column-major (count, nx)
        _lib.check(self._L.tiktak_cuda_sobol_points(self._h, first, count, out.ctypes.data), "setupSobol")
        return out.T

    def objFun(self, theta):
        theta = np.atleast_2d(np.asarray(theta, dtype=np.float64))
        n = theta.shape[0]
        x = np.ascontiguousarray(theta.T)  # column-major (n, nx)
        f = np.empty(n, dtype=np.float64)
        _lib.check(self._L.tiktak_cuda_objfun(self._h, x.ctypes.data, n, f.ctypes.data), "objFun")
        return f

    # ------------------------------------------------------------------ stages 3-5
    def solveAtSobolPoints(self):
        """Returns (complete, n_evaluated, n_legit) -- `complete` as in :465."""
        ne, nl = C.c_int64(), C.c_int64()
        if self.world == 1:
            _lib.check(self._L.tiktak_cuda_pretest(self._h, C.byref(ne), C.byref(nl)), "solveAtSobolPoints")
            self.n_evaluated, self.n_legit = ne.value, nl.value
            return True, ne.value, nl.value
        return self._solve_sharded()

    def _solve_sharded(self):
        from . import dist as D

        N, Lg = self.cfg.qr_ndraw, self.cfg.legitimate
        nblocks = N // SOBOL_BLOCK + 1
        can_cut = Lg <= N
        wave = 64 * self.world if can_cut else nblocks
        dev = f"cuda:{self.device}"
        cum, kcut, nleg = 0, -1, 0
        cb0 = 0
        if N == 0 or Lg <= 0:
            _lib.check(self._L.tiktak_cuda_pretest_wave(self._h, 0, 0), "pretest_wave")
            kcut, nleg, cb0 = 0, 0, nblocks
        while cb0 < nblocks and kcut < 0:
            ncb = min(wave, nblocks - cb0)
            _lib.check(self._L.tiktak_cuda_pretest_wave(self._h, cb0, ncb), "pretest_wave")
            counts = np.zeros(nblocks, dtype=np.int64)
            _lib.check(self._L.tiktak_cuda_pretest_counts(self._h, counts.ctypes.data, nblocks), "pretest_counts")
            counts[cb0 + ncb:] = 0
            # one exchange per wave: all-reduce of the per-block valid counts
            _, b, need, cum = D.global_cut(counts, cb0, Lg if can_cut else 2**62, cum, self.group, dev)
            if b >= 0:
                n = C.c_int64()
                _lib.check(self._L.tiktak_cuda_pretest_find(self._h, b, need, C.byref(n)), "pretest_find")
                kcut, nleg = D.agree_max(n.value, self.group, dev), Lg
            cb0 += ncb
        if kcut < 0:
            kcut, nleg = N, cum
        _lib.check(self._L.tiktak_cuda_pretest_set_cut(self._h, kcut, nleg), "pretest_set_cut")
        self.n_evaluated, self.n_legit = kcut, nleg
        return True, kcut, nleg

    def sobolFnVal(self, first=1, count=None, out=None):
        """fval column of sobolFnVal.dat rows first.. (:467); NaN for points of other shards / beyond the cut.
        `out`: optional preallocated float64 array (e.g. a view of pinned host memory)."""
        count = self.n_evaluated if count is None else count
        f = np.empty(count, dtype=np.float64) if out is None else out[:count]
        _lib.check(self._L.tiktak_cuda_pretest_values(self._h, first, count, f.ctypes.data), "sobolFnVal")
        return f

    # ------------------------------------------------------------------ stage 6
    def chooseSobol(self):
        K, nx = self.K, self.cfg.nx
        xs = np.empty((nx + 1, K), dtype=np.float64)
        idx = np.zeros(K, dtype=np.int32)
        _lib.check(self._L.tiktak_cuda_choose(self._h, xs.ctypes.data, idx.ctypes.data), "chooseSobol")
        if self.world > 1 and K > 1:
            from . import dist as D

            M = K - 1
            cand = np.empty((M, 2), dtype=np.float64)
            _lib.check(self._L.tiktak_cuda_choose_candidates(self._h, cand.ctypes.data), "choose_candidates")
            allh = D.gather_candidates(cand, self.world, self.group, f"cuda:{self.device}")
            _lib.check(self._L.tiktak_cuda_choose_merge(self._h, allh.ctypes.data, allh.shape[0], xs.ctypes.data,
                                                        idx.ctypes.data), "choose_merge")
        self.x_starts = xs.T.copy()  # (K, nx+1) = [fval, x]
        self.sobol_idx = idx
        return self.x_starts

    # ------------------------------------------------------------------ stage 7
    def rounds(self):
        B = max(1, int(self.cfg.round_size))
        k = 1
        while k <= self.K:
            n = min(B, self.K - k + 1)
            yield k, n
            k += n

    def LocalMinimizations(self, algor="a"):
        """Runs every restart; returns `complete` (:536).  Rows land in searchStart / searchResults."""
        alg = _ALG[algor]
        K, nx = self.K, self.cfg.nx
        starts = np.full((K, nx), np.nan)
        results = np.full((K, nx + 4), np.nan)
        evals = np.zeros(K, dtype=np.int32)
        for first_k, n_k in self.rounds():
            s = np.empty((nx, n_k), dtype=np.float64)
            r = np.empty((nx + 4, n_k), dtype=np.float64)
            e = np.zeros(n_k, dtype=np.int32)
            _lib.check(self._L.tiktak_cuda_local(self._h, alg, first_k, n_k, s.ctypes.data, r.ctypes.data, e.ctypes.data),
                       "LocalMinimizations")
            if self.world > 1:
                r, e = self._exchange_round(r, e, n_k)
            sl = slice(first_k - 1, first_k - 1 + n_k)
            starts[sl] = s.T
            results[sl] = r.T
            evals[sl] = e
        self.searchStart, self.searchResults, self.evals = starts, results, evals
        return True

    def _exchange_round(self, r, e, n_k):
        """one all-gather of this round's rows over NVLink; every rank then ingests all of them."""
        from . import dist as D

        nx = self.cfg.nx
        pack = np.concatenate([r, e[None, :].astype(np.float64)], axis=0)  # (nx+5, n_k)
        merged = D.exchange_rows(pack, self.world, self.group, f"cuda:{self.device}")
        rows = np.ascontiguousarray(merged[: nx + 4])
        _lib.check(self._L.tiktak_cuda_push_results(self._h, rows.ctypes.data, n_k), "push_results")
        # the #improvements column is assigned at ingestion (identically on every rank)
        out = rows.copy()
        out[2] = self._improvements(rows)
        return out, merged[nx + 4].astype(np.int32)

    def _improvements(self, rows):
        if not hasattr(self, "_bestf"):
            self._bestf, self._bestimp = None, 0
        imp = np.zeros(rows.shape[1])
        for s in range(rows.shape[1]):
            f = rows[3, s]
            if self._bestf is None:
                self._bestf = self.x_starts[0, 0]
            if f < self._bestf:
                self._bestimp += 1
                self._bestf = f
            imp[s] = self._bestimp
        return imp

    # ------------------------------------------------------------------ queries
    def getBestLine(self):
        f, k = C.c_double(), C.c_int32()
        x = np.empty(self.cfg.nx, dtype=np.float64)
        _lib.check(self._L.tiktak_cuda_best(self._h, C.byref(f), dptr(x), C.byref(k)), "getBestLine")
        return f.value, x, k.value

    def lastSearch(self):
        """state 10 (:635-681): BFGS polish (EST_dfpmin) from the best row; returns (fn_val, evalParam, iterations)."""
        f, it = C.c_double(), C.c_int32()
        x = np.empty(self.cfg.nx, dtype=np.float64)
        _lib.check(self._L.tiktak_cuda_last_search(self._h, C.byref(f), dptr(x), C.byref(it)), "lastSearch")
        return f.value, x, it.value

    # ------------------------------------------------------------------ the single-instance state machine
    def run(self, algor="a"):
        """Cold start, states 2..8 (TiktakGlobalSearch.f90:217-401) with nothing missing."""
        complete, ne, nl = self.solveAtSobolPoints()
        self.chooseSobol()
        self.LocalMinimizations(algor)
        return self.getBestLine()
