"""CPU tests of the N>1 path: the three exchanges of tiktak_b200/dist.py over gloo, world_size 2.

The per-rank partial data a GPU shard would produce (valid counts of the count blocks it owns, its
local top-M candidates, the result rows of the restarts dealt to it) is produced here with the
oracle; the test checks that the sharded protocol reconstructs exactly the single-rank answer.
"""
import os
import socket

import numpy as np
import pytest

from tiktak_b200 import OBJ_GRIEWANK, SOBOL_BLOCK, make_config


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, q):
    import torch.distributed as dist

    from oracle.oracle import MATH_TK, Oracle
    from tiktak_b200 import dist as D

    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        N, M = 3 * SOBOL_BLOCK + 1234, 25
        Lg = int(0.35 * N)  # about 70 % of the valid points (valid = below the median)
        cfg = make_config(10, 1, N, M, -400.0, 600.0, objective_id=OBJ_GRIEWANK)
        full = Oracle(cfg, MATH_TK); full.solveAtSobolPoints()
        f = full.sobolFnVal()  # index n-1
        med = float(np.median(f))
        valid = f < med
        nblocks = N // SOBOL_BLOCK + 1
        blk = (np.arange(1, N + 1)) // SOBOL_BLOCK  # count block of point n (n = 0 is the skipped vector)
        # --- 1. legitimate cut from sharded counts
        counts = np.zeros(nblocks, dtype=np.int64)
        for b in range(rank, nblocks, world):
            counts[b] = int(valid[blk == b].sum())
        cg, b, need, cum = D.global_cut(counts, 0, Lg, 0)
        cums = np.cumsum(valid)
        kcut_true = int(np.argmax(cums >= Lg)) + 1
        assert b == kcut_true // SOBOL_BLOCK and cum + need == Lg
        mine = 0
        if b % world == rank:  # owner finds the exact index inside its block
            inblk = np.nonzero((blk == b) & valid)[0]
            mine = int(inblk[need - 1]) + 1
        kcut = D.agree_max(mine)
        assert kcut == kcut_true
        # --- 2. candidates: local top-M of owned, evaluated points (+ trailer row) -> merge == global top-M
        import torch

        own = np.nonzero((blk % world == rank) & (np.arange(1, N + 1) <= kcut))[0]
        order = own[np.lexsort((own, f[own]))][:M]
        cand = np.full((M + 1, 2), np.nan); cand[: len(order), 0] = f[order]; cand[: len(order), 1] = order + 1
        cand[M] = [int(valid[own].sum()), len(order)]  # trailer: [#valid points of the shard, #rows filled]
        allc = D.gather_candidates(torch.as_tensor(cand), world).numpy().reshape(world, M + 1, 2)
        rows = np.concatenate([allc[r, : int(allc[r, M, 1])] for r in range(world)])
        merged = rows[np.lexsort((rows[:, 1], rows[:, 0]))][:M]
        ev = np.arange(kcut)
        glob = ev[np.lexsort((ev, f[:kcut]))][:M]
        assert np.array_equal(merged[:, 1].astype(int), glob + 1)
        assert int(allc[:, M, 0].sum()) == int(valid[:kcut].sum())
        # --- 3. restart rounds: slot s is run by rank s % world and travels as row s // world of that rank's pack
        for n_k in (7, 8, 1):
            ncol = 10 + 2
            truth = np.arange(ncol * n_k, dtype=np.float64).reshape(n_k, ncol) + 0.25
            mine = np.full((n_k, ncol), np.nan)
            mine[rank::world] = truth[rank::world]  # this rank only knows its own slots
            send = torch.as_tensor(D.pack_rows_ref(mine, rank, world))
            recv = torch.empty((world * send.shape[0], ncol), dtype=torch.float64)
            D.gather_rows(recv, send)
            assert np.array_equal(D.unpack_rows_ref(recv.numpy(), n_k, world), truth)
        fmin, kmin = D.best_allreduce(3.0 - rank, 10 + rank)
        assert (fmin, kmin) == (3.0 - (world - 1), 10 + world - 1)
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        import traceback
        q.put((rank, traceback.format_exc()))
    finally:
        dist.destroy_process_group()


def test_sharded_protocol_world2(built):
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res
