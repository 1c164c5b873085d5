"""Codeword-level data parallelism: Monte-Carlo codewords are independent, so global
codeword index g goes to rank g mod world, every rank keeps a replica of the (read-only) operator and graph
tables, and the only collective is a sum of error / iteration counters (NCCL over NVLink on GPUs, gloo in the
CPU tests).  No data-path collective exists or is invented."""
import numpy as np
import torch
import torch.distributed as dist


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_indices(total, rank=None, world_size=None):
    """Global codeword indices decoded by this rank: g = rank, rank + world, ... (reference draw order is
    preserved inside every rank because all ranks walk the same host RNG stream)."""
    if rank is None:
        rank, world_size = world()
    return np.arange(rank, total, world_size)


def allreduce_counts(counts, group=None):
    """Sum int64 counters [..] over ranks; returns a host numpy array.  Works on CUDA (nccl) or CPU (gloo)."""
    t = counts if torch.is_tensor(counts) else torch.as_tensor(np.asarray(counts))
    t = t.to(torch.int64)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        if dist.get_backend(group) == "nccl" and not t.is_cuda:
            t = t.cuda()
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t.cpu().numpy()


def gather_rows(rows, total, group=None):
    """All-gather per-codeword result rows (float64 [n_local, k]) back into global order [total, k] so that the
    sequential stop rule (sparc_ldpc.py:1217-1245) can be replayed identically on every rank."""
    rank, ws = world()
    rows = np.asarray(rows, dtype=np.float64)
    k = rows.shape[1] if rows.ndim == 2 else 1
    out = np.zeros((total, k))
    mine = shard_indices(total, rank, ws)
    out[mine] = rows.reshape(len(mine), k)
    if ws > 1:
        t = torch.from_numpy(out)
        if dist.get_backend(group) == "nccl":
            t = t.cuda()
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)   # disjoint supports: sum == gather
        out = t.cpu().numpy()
    return out


def replay_stop_rule(error_flags, min_errors, max_blocks):
    """Number of blocks the reference's `while nblockerrors < MIN_ERRORS` loop would have used, given the
    per-block error flags in draw order (sparc_ldpc.py:1217-1245)."""
    nerr = 0
    for i, e in enumerate(error_flags):
        nerr += 1 if e else 0
        if nerr >= min_errors or i + 1 >= max_blocks:
            return i + 1
    return len(error_flags)
