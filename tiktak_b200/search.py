"""Host-side mirror of the reference's stage procedures over the C ABI.

The method names, argument meaning and error behaviour follow the internal procedures of
PROGRAM TiktakGlobalSearch (TiktakGlobalSearch.f90:404-1148) for the hot path:

    setupSobol            :825-838     sobol_trial(i,:) = lo + q*(hi-lo)
    solveAtSobolPoints    :405-471     objFun at points 1.. until `legitimate` valid values
    chooseSobol           :840-899     x_starts = [p_init; best maxpoints Sobol points]
    LocalMinimizations    :473-544     restarts k = 1..p_maxpoints (getModifiedParam + completeSearch)
    getBestLine           :781-802     best searchResults row so far
    run                   :217-401     the state machine 1 -> 2 -> 3 -> 6 -> 7 (-> 10) of one cold start

All compute happens in libtiktak_cuda (CUDA, sm_100a).  With world > 1 the Sobol index range and each
launch's restarts are sharded over the ranks; the exchanges (valid counts, candidate rows, result rows)
are NCCL collectives issued by the library itself on the handle's communicator -- torch.distributed only
carries the communicator's 128-byte id from rank 0 to the other ranks when the instance is created.
"""
import ctypes as C

import numpy as np

from . import lib as _lib
from ._abi import ALG_AMOEBA, ALG_DFNLS, SOBOL_BLOCK, TiktakConfig, as_f64, dptr, smm_user
from .config import TikTakConfig

_ALG = {"a": ALG_AMOEBA, "b": ALG_DFNLS, "d": 2, "amoeba": ALG_AMOEBA, "dfnls": ALG_DFNLS, "dfpmin": 2, 0: ALG_DFNLS, 1: ALG_AMOEBA, 2: 2}


class TikTakSearch:
    """One "instance" of the reference program bound to one GPU (rank of a torch.distributed group or alone)."""

    def __init__(self, cfg: TikTakConfig, device=0, rank=0, world=1, group=None, comm=True):
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
        if world > 1 and comm:  # comm=False: the caller drives the exchanges itself (enqueue / gather / ingest of the C ABI)
            self._init_comm()
        self.n_evaluated = None
        self._n_legit = None
        self.x_starts = None
        self.sobol_idx = None
        self.searchStart = None  # rows of searchStart.dat: (K, nx)
        self.searchResults = None  # rows of searchResults.dat: (K, nx+4) = seqNo, k, #improv, fval, x
        self.evals = None

    def _init_comm(self):
        """the handle's NCCL communicator: rank 0 creates the id, the process group carries it to the other ranks"""
        import torch.distributed as dist

        buf = C.create_string_buffer(128)
        if self.rank == 0:
            _lib.check(self._L.tiktak_cuda_comm_unique_id(buf, 128), "comm_unique_id")
        box = [bytes(buf.raw)]
        dist.broadcast_object_list(box, src=0, group=self.group)
        ident = C.create_string_buffer(box[0], 128)
        _lib.check(self._L.tiktak_cuda_comm_init(self._h, ident, 128), "comm_init")

    @property
    def n_legit(self):
        """#(fval < fvalmax) among the evaluated points (the final legitSobol.dat value); read back from the device on
        first use when the stage did not need it to place the cut"""
        if self._n_legit is None and self.n_evaluated is not None:
            self._resolve_legit()
        return self._n_legit

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

    def stream_ptr(self):
        """the cudaStream_t every kernel of this handle is queued on"""
        return int(self._L.tiktak_cuda_stream(self._h) or 0)

    def torch_stream(self):
        """that stream as a torch stream: collectives issued under `with torch.cuda.stream(...)` are ordered with
        the library's kernels on the device, without host synchronisation"""
        import torch

        if getattr(self, "_ext_stream", None) is None:
            self._ext_stream = torch.cuda.ExternalStream(self.stream_ptr(), device=torch.device("cuda", self.device))
        return self._ext_stream

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
    def solveAtSobolPoints(self, wait=True):
        """Returns (complete, n_evaluated, n_legit) -- `complete` as in :465.

        wait=False: when the `legitimate` cut cannot fire (legitimate > qr_ndraw) the evaluation is only queued on
        the device and n_legit comes back as None; the `n_legit` property reads it back on first use."""
        out = self._solve(wait)
        if wait and self._n_legit is None:
            out = (out[0], out[1], self.n_legit)
        return out

    def _resolve_legit(self):
        nl = C.c_int64()  # world > 1 and not known yet: a collective over the handle's communicator (every rank calls it)
        _lib.check(self._L.tiktak_cuda_pretest_result(self._h, None, C.byref(nl)), "pretest_result")
        self._n_legit = nl.value

    def _solve(self, wait):
        ne, nl = C.c_int64(), C.c_int64()
        if not wait and self.cfg.legitimate > self.cfg.qr_ndraw > 0:
            # the cut cannot fire: the evaluation is only queued; the valid count is read back by chooseSobol / n_legit
            _lib.check(self._L.tiktak_cuda_pretest(self._h, C.byref(ne), None), "solveAtSobolPoints")
            self.n_evaluated, self._n_legit = ne.value, None
            return True, ne.value, None
        _lib.check(self._L.tiktak_cuda_pretest(self._h, C.byref(ne), C.byref(nl)), "solveAtSobolPoints")
        self.n_evaluated, self._n_legit = ne.value, nl.value
        return True, ne.value, nl.value

    def sobolFnVal(self, first=1, count=None, out=None):
        """fval column of sobolFnVal.dat rows first.. (:467); NaN for points of other shards / beyond the cut.
        `out`: optional preallocated float64 array (e.g. a view of pinned host memory)."""
        count = self.n_evaluated if count is None else count
        f = np.empty(count, dtype=np.float64) if out is None else out[:count]
        _lib.check(self._L.tiktak_cuda_pretest_values(self._h, first, count, f.ctypes.data), "sobolFnVal")
        return f

    def sobolFnValShard(self, out=None):
        """this shard's fval store in one copy (world > 1): count blocks rank, rank+world, ... of SOBOL_BLOCK values"""
        cap = (self.cfg.qr_ndraw // SOBOL_BLOCK + 1 + self.world - 1) // self.world * SOBOL_BLOCK
        f = np.empty(cap, dtype=np.float64) if out is None else out
        n = C.c_int64()
        _lib.check(self._L.tiktak_cuda_pretest_values_shard(self._h, f.ctypes.data, f.size, C.byref(n)), "sobolFnValShard")
        return f[: n.value]

    def sobolFnValShardAsync(self, out):
        """queue the copy of this shard's stored values into `out` (pinned host memory) on a second stream; returns the number
        of values; sobolFnValWait() returns when they have landed.  Nothing on the path waits for them (they are
        sobolFnVal.dat's fval column), so the copy runs beside selection and the restarts."""
        n = C.c_int64()
        _lib.check(self._L.tiktak_cuda_pretest_values_shard_async(self._h, out.ctypes.data, out.size, C.byref(n)), "sobolFnValShardAsync")
        return n.value

    def sobolFnValWait(self):
        _lib.check(self._L.tiktak_cuda_pretest_values_wait(self._h), "sobolFnValWait")

    # ------------------------------------------------------------------ stage 6
    def chooseSobol(self):
        K, nx = self.K, self.cfg.nx
        xs = np.empty((nx + 1, K), dtype=np.float64)
        idx = np.zeros(K, dtype=np.int32)
        # world > 1: shard-local top-(K-1), one all-gather of the candidate rows inside the library, merged on the device
        _lib.check(self._L.tiktak_cuda_choose(self._h, xs.ctypes.data, idx.ctypes.data), "chooseSobol")
        if self._n_legit is None and self.world > 1:  # no-cut run: the global valid count travelled in the trailer rows
            self._resolve_legit()
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

    def LocalMinimizations(self, algor="a", window=None):
        """Runs every restart; returns `complete` (:536).  Rows land in searchStart / searchResults.

        Rounds of cfg.round_size restarts are blended against the best row at the round's start (DESIGN.md "round
        schedule").  A launch holds `window` restarts = several rounds: the later ones are started early against the
        same best row and kept only while that row stays the best (tiktak_cuda_local_ingest_rounds), so the accepted
        sequence is exactly that of one round after the other while the idle SMs of a small round do useful work.
        window=None: one resident wave of the local-search kernel (tiktak_cuda_local_capacity); window=0: no
        early starts.  Per launch: blend -> local searches -> [one all-gather of the rows over NCCL, world > 1] ->
        ingest; the host reads one integer per launch and all rows once at the end."""
        alg = _ALG[algor]
        K, nx, W = self.K, self.cfg.nx, self.world
        B = max(1, int(self.cfg.round_size))
        if window is None:
            cap = C.c_int32()
            _lib.check(self._L.tiktak_cuda_local_capacity(self._h, alg, C.byref(cap)), "local_capacity")
            window = cap.value  # 0 = the mode forbids early starts: one round per launch
        per_launch = max(B, (int(window) // B) * B)  # whole rounds
        k, acc = 1, C.c_int32()
        self.local_launches = 0
        while k <= K:
            n = min(per_launch, K - k + 1)
            _lib.check(self._L.tiktak_cuda_local_launch(self._h, alg, k, n, B, C.byref(acc)), "LocalMinimizations")
            assert 0 < acc.value <= n
            k += acc.value
            self.local_launches += 1
        s = np.empty((nx, K), dtype=np.float64)
        r = np.empty((nx + 4, K), dtype=np.float64)
        e = np.zeros(K, dtype=np.int32)
        _lib.check(self._L.tiktak_cuda_local_fetch(self._h, 1, K, s.ctypes.data, r.ctypes.data, e.ctypes.data),
                   "LocalMinimizations")
        self.searchStart, self.searchResults, self.evals = s.T, r.T, e  # (K, nx) / (K, nx+4) views of the column-major buffers: no host transposition
        st = np.zeros(K, dtype=np.int32)
        _lib.check(self._L.tiktak_cuda_local_status(self._h, 1, K, st.ctypes.data_as(C.POINTER(C.c_int32))), "local_status")
        self.status = st  # 0 = ended like the reference's search; 2 = DFNLS refused the start
        nr = np.zeros(K, dtype=np.int32)
        _lib.check(self._L.tiktak_cuda_local_nrescue(self._h, 1, K, nr.ctypes.data_as(C.POINTER(C.c_int32))), "local_nrescue")
        self.nrescue = nr  # RESCUE_H calls per restart (DFNLS)
        return True

    def asyncSupported(self, algor="a"):
        yes = C.c_int32()
        _lib.check(self._L.tiktak_cuda_local_async_supported(self._h, _ALG[algor], C.byref(yes)), "local_async_supported")
        return bool(yes.value)

    def asyncClock(self, algor="a", instances=None):
        """the virtual clock of LocalMinimizationsAsync(algor, instances): 0 = a restart costs PICKUP + c0 + trips + shrinks*c0 ticks
        (four-warp instances); 1 = PICKUP + its evaluations (one-warp instances, DFNLS)"""
        c = C.c_int32()
        P = max(1, int(self.cfg.round_size if instances is None else instances))
        _lib.check(self._L.tiktak_cuda_local_async_clock(self._h, _ALG[algor], P, C.byref(c)), "local_async_clock")
        return c.value

    def LocalMinimizationsAsync(self, algor="a", instances=None, free_running=False):
        """Every restart by `instances` asynchronous instances (default: cfg.round_size) -- the reference's multi-process mode on
        one GPU, one kernel launch (tiktak_cuda_local_async; DESIGN.md "asynchronous instances").  Sets searchStart,
        searchResults, evals and `instances_used`."""
        alg = _ALG[algor]
        K, nx = self.K, self.cfg.nx
        P = max(1, int(self.cfg.round_size if instances is None else instances))
        used = C.c_int32()
        _lib.check(self._L.tiktak_cuda_local_async(self._h, alg, 1, K, -P if free_running else P, C.byref(used)), "LocalMinimizationsAsync")
        self.instances_used = used.value
        base, order = np.zeros(K, dtype=np.int32), np.zeros(K, dtype=np.int32)
        _lib.check(self._L.tiktak_cuda_local_async_info(self._h, 1, K, base.ctypes.data_as(C.POINTER(C.c_int32)), order.ctypes.data_as(C.POINTER(C.c_int32))),
                   "local_async_info")
        self.async_base, self.async_order = base, order  # the row each start was blended toward (-1: none); place in the file order
        self.local_launches = 1
        s = np.empty((nx, K), dtype=np.float64)
        r = np.empty((nx + 4, K), dtype=np.float64)
        e = np.zeros(K, dtype=np.int32)
        _lib.check(self._L.tiktak_cuda_local_fetch(self._h, 1, K, s.ctypes.data, r.ctypes.data, e.ctypes.data), "LocalMinimizationsAsync")
        self.searchStart, self.searchResults, self.evals = s.T, r.T, e  # (K, nx) / (K, nx+4) views of the column-major buffers: no host transposition
        st = np.zeros(K, dtype=np.int32)
        _lib.check(self._L.tiktak_cuda_local_status(self._h, 1, K, st.ctypes.data_as(C.POINTER(C.c_int32))), "local_status")
        nr = np.zeros(K, dtype=np.int32)
        _lib.check(self._L.tiktak_cuda_local_nrescue(self._h, 1, K, nr.ctypes.data_as(C.POINTER(C.c_int32))), "local_nrescue")
        self.status, self.nrescue = st, nr
        return True

    def pushResults(self, rows):
        """rows (n, nx+4) = seqNo, k, #improvements, fval, x re-read from searchResults.dat (options 3, missing-work sweep)"""
        rows = np.asarray(rows, dtype=np.float64)
        rin = np.ascontiguousarray(rows.T)
        _lib.check(self._L.tiktak_cuda_push_results(self._h, rin.ctypes.data, len(rows)), "pushResults")

    def SolveMissingSearch(self, algor, missing):
        """states 8-9 (:1048-1146): the restarts `missing` (never reported by a killed instance), one after the other; each
        starts from getModifiedParam(p_maxpoints, x_starts(k,2:)) and ends in completeSearch(alg, k).  Returns the rows."""
        alg, nx = _ALG[algor], self.cfg.nx
        out = []
        for k in missing:
            st, row, ev = np.empty(nx), np.empty(nx + 4), C.c_int32()
            _lib.check(self._L.tiktak_cuda_local_missing(self._h, alg, int(k), dptr(st), dptr(row), C.byref(ev)), "SolveMissingSearch")
            out.append((st, row, ev.value))
        return out

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
