"""The exchanges of the sharded run (one process per GPU, torch.distributed; NCCL over NVLink on the GPU box).

Every exchange works on tensors that already live where the collective runs (device memory of the rank's GPU
under NCCL, host memory under gloo in the CPU tests), so on the GPU the host only queues work:
  * reduce_counts     -- all-reduce of the per-block valid counts -> the `legitimate` cut (find_cut)
  * gather_candidates -- all-gather of each rank's local top-(K-1) list [fval, idx] + trailer row
  * gather_rows       -- all-gather of one restart round's packed result rows; slot s of the round is run by
                         rank s % world and travels as row s // world of that rank's pack
The layout of the packed rows is fixed by the device kernels (pack_kernel / ingest_kernel in
csrc/kernels_nm.cu); pack_rows_ref / unpack_rows_ref restate it in numpy for the CPU tests.
"""
import numpy as np


def _torch():
    import torch
    import torch.distributed as dist

    return torch, dist


def reduce_counts(counts_t, group=None):
    """in-place SUM all-reduce of an int64 tensor of per-count-block valid counts (zeros for other ranks' blocks)"""
    _, dist = _torch()
    dist.all_reduce(counts_t, op=dist.ReduceOp.SUM, group=group)
    return counts_t


def find_cut(counts_global, first_block, legitimate, cum_before):
    """first block b >= first_block whose cumulative valid count reaches `legitimate` -> (b or -1, valid points
    still needed inside b, cumulative count before b)"""
    cum = int(cum_before)
    for b in range(int(first_block), len(counts_global)):
        c = int(counts_global[b])
        if cum + c >= legitimate:
            return b, legitimate - cum, cum
        cum += c
    return -1, 0, cum


def global_cut(counts_local, first_block, legitimate, cum_before, group=None, device="cpu"):
    """counts_local: int64 array over ALL count blocks (zeros for blocks of other ranks).
    Returns (counts_global, cut_block or -1, need, cum): see find_cut."""
    torch, _ = _torch()
    t = torch.as_tensor(np.asarray(counts_local, dtype=np.int64)).to(device)
    cg = reduce_counts(t, group).cpu().numpy()
    b, need, cum = find_cut(cg, first_block, legitimate, cum_before)
    return cg, b, need, cum


def agree_max(value, group=None, device="cpu"):
    torch, dist = _torch()
    t = torch.tensor([int(value)], dtype=torch.int64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return int(t.item())


def gather_candidates(cand_t, world, group=None):
    """cand_t (K, 2) tensor [fval, idx] rows + trailer -> (world*K, 2) tensor on the same device"""
    torch, dist = _torch()
    allc = torch.empty((world * cand_t.shape[0], cand_t.shape[1]), dtype=cand_t.dtype, device=cand_t.device)
    dist.all_gather_into_tensor(allc, cand_t, group=group)
    return allc


def gather_rows(recv_t, send_t, group=None):
    """one all-gather per restart round: recv = [rank 0 pack | rank 1 pack | ...]"""
    _, dist = _torch()
    dist.all_gather_into_tensor(recv_t, send_t, group=group)
    return recv_t


def n_max(n_k, world):
    return (n_k + world - 1) // world


def pack_rows_ref(rows, rank, world):
    """rows (n_k, ncol) of the whole round (only this rank's slots need to be valid) -> this rank's pack
    (n_max, ncol): row q = slot rank + q*world.  numpy restatement of pack_kernel."""
    n_k, ncol = rows.shape
    out = np.zeros((n_max(n_k, world), ncol), dtype=rows.dtype)
    own = np.arange(rank, n_k, world)
    out[: len(own)] = rows[own]
    return out


def unpack_rows_ref(gathered, n_k, world):
    """gathered (world*n_max, ncol) -> (n_k, ncol): slot s = row s // world of rank s % world (ingest_kernel)"""
    nm = n_max(n_k, world)
    g = gathered.reshape(world, nm, -1)
    s = np.arange(n_k)
    return g[s % world, s // world]


def best_allreduce(fval, k, group=None, device="cpu"):
    """argmin all-reduce of (fval, k): returns the global (fmin, kmin); ties -> lowest k."""
    torch, dist = _torch()
    f = torch.tensor([float(fval)], dtype=torch.float64, device=device)
    dist.all_reduce(f, op=dist.ReduceOp.MIN, group=group)
    kk = torch.tensor([int(k) if float(fval) == float(f.item()) else 2**62], dtype=torch.int64, device=device)
    dist.all_reduce(kk, op=dist.ReduceOp.MIN, group=group)
    return float(f.item()), int(kk.item())
