"""The three exchanges of the sharded run (one process per GPU, torch.distributed).

Kept free of CUDA calls so the same code runs over NCCL on the GPU box and over gloo in the CPU tests:
  * global_cut       -- all-reduce of the per-block valid counts -> the `legitimate` cut (which block,
                        how many valid points are still needed inside it)
  * gather_candidates-- all-gather of each rank's local top-(K-1) [fval, idx] rows
  * exchange_rows    -- all-gather of one restart round's result rows; slot s was run by rank s % world
"""
import numpy as np


def _torch():
    import torch
    import torch.distributed as dist

    return torch, dist


def global_cut(counts_local, first_block, legitimate, cum_before, group=None, device="cpu"):
    """counts_local: int64 array over ALL count blocks (zeros for blocks of other ranks).
    Returns (counts_global, cut_block or -1, need, cum_after): the first block b >= first_block whose
    cumulative valid count reaches `legitimate`, and how many valid points inside b are still needed."""
    torch, dist = _torch()
    t = torch.as_tensor(np.asarray(counts_local, dtype=np.int64)).to(device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    cg = t.cpu().numpy()
    cum = cum_before
    for b in range(first_block, len(cg)):
        c = int(cg[b])
        if cum + c >= legitimate:
            return cg, b, legitimate - cum, cum
        cum += c
    return cg, -1, 0, cum


def agree_max(value, group=None, device="cpu"):
    torch, dist = _torch()
    t = torch.tensor([int(value)], dtype=torch.int64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return int(t.item())


def gather_candidates(cand_local, world, group=None, device="cpu"):
    """cand_local (M, 2) [fval, idx] (NaN rows = absent) -> (world*M, 2) on the host."""
    torch, dist = _torch()
    mine = torch.as_tensor(np.ascontiguousarray(cand_local, dtype=np.float64)).to(device)
    allc = torch.empty((world * mine.shape[0], 2), dtype=torch.float64, device=device)
    dist.all_gather_into_tensor(allc, mine, group=group)
    return np.ascontiguousarray(allc.cpu().numpy())


def exchange_rows(pack_local, world, group=None, device="cpu"):
    """pack_local (ncol, n_k): this rank's columns are valid for slots s with s % world == rank."""
    torch, dist = _torch()
    mine = torch.as_tensor(np.ascontiguousarray(pack_local, dtype=np.float64)).to(device)
    ncol, n_k = mine.shape
    allr = torch.empty((world * ncol, n_k), dtype=torch.float64, device=device)  # rank-major concatenation
    dist.all_gather_into_tensor(allr, mine, group=group)
    allh = allr.cpu().numpy().reshape(world, ncol, n_k)
    merged = np.empty_like(allh[0])
    for s in range(merged.shape[1]):
        merged[:, s] = allh[s % world, :, s]
    return merged


def best_allreduce(fval, k, group=None, device="cpu"):
    """argmin all-reduce of (fval, k): returns the global (fmin, kmin); ties -> lowest k."""
    torch, dist = _torch()
    f = torch.tensor([float(fval)], dtype=torch.float64, device=device)
    dist.all_reduce(f, op=dist.ReduceOp.MIN, group=group)
    kk = torch.tensor([int(k) if float(fval) == float(f.item()) else 2**62], dtype=torch.int64, device=device)
    dist.all_reduce(kk, op=dist.ReduceOp.MIN, group=group)
    return float(f.item()), int(kk.item())
