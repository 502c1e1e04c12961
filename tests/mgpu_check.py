"""torchrun --nproc-per-node N tests/mgpu_check.py : sharded run vs the single-instance oracle (rank 0 checks).

A multi-process GPU parity test (not collected by pytest: it needs torchrun and N GPUs); it lives under tests/ because
it imports the oracle, which only test code may do."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.oracle import MATH_TK, Oracle  # noqa: E402
from tiktak_b200 import OBJ_GRIEWANK, OBJ_RASTRIGIN, SOBOL_BLOCK, TikTakSearch, make_config  # noqa: E402

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
ok = True
for name, kw in {
    "griewank10_nocut": dict(nx=10, nm=1, qr_ndraw=5 * SOBOL_BLOCK + 777, maxpoints=63, lo=-400.0, hi=600.0, init=[100.0] * 10,
                             objective_id=OBJ_GRIEWANK, round_size=16),
    "rastrigin20_cut": dict(nx=20, nm=1, qr_ndraw=9 * SOBOL_BLOCK + 5, maxpoints=40, lo=-4.12, hi=5.12, init=[1.0] * 20,
                            objective_id=OBJ_RASTRIGIN, round_size=7, fvalmax=140000.0, legitimate=int(0.34 * (9 * SOBOL_BLOCK + 5))),
}.items():
    cfg = make_config(**kw)
    with TikTakSearch(cfg, device=lr, rank=rank, world=world) as s:
        _, ne, _ = s.solveAtSobolPoints()
        xs = s.chooseSobol()
        nl = s.n_legit  # without a possible cut the global valid count arrives with the candidate all-gather
        s.LocalMinimizations("a")
        res, st, ev, best = s.searchResults, s.searchStart, s.evals, s.getBestLine()
    if rank == 0:
        o = Oracle(cfg, MATH_TK)
        _, one, onl = o.solveAtSobolPoints()
        oxs = o.chooseSobol(mode=1)
        o.LocalMinimizations(1)
        checks = {"cut": (ne, nl) == (one, onl), "x_starts": np.array_equal(xs, oxs), "starts": np.array_equal(st, o.searchStart),
                  "results": np.array_equal(res[:, [1, 3] + list(range(4, res.shape[1]))], o.searchResults[:, [1, 3] + list(range(4, res.shape[1]))]),
                  "imp": np.array_equal(res[:, 2], o.searchResults[:, 2]), "evals": np.array_equal(ev, o.evals),
                  "best": best[0] == o.getBestLine()[0]}
        print(name, "world", world, "ne/nl", ne, nl, one, onl, checks, flush=True)
        ok = ok and all(checks.values())
dist.barrier()
dist.destroy_process_group()
if rank == 0:
    print("MGPU_CHECK", "PASS" if ok else "FAIL")
    sys.exit(0 if ok else 1)
