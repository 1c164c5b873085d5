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


# ---------------------------------------------------------------------------- sharded chunk decode (parity mode)
def group_info(group):
    """(rank, world) of `group`; (0, 1) for group=None (single-rank drivers never touch torch.distributed)."""
    if group is None:
        return 0, 1
    g = None if group is True else group           # True = the default (world) group
    return dist.get_rank(g), dist.get_world_size(g)


def _flatten_row(row):
    """A result row (tuple / list of scalars and 1-D arrays) -> float64 vector [n_elems, len_0 (-1 = scalar), ...,
    values...] so that rows of any flow travel through one all-reduce."""
    parts, lens = [], []
    items = row if isinstance(row, (tuple, list)) else (row,)
    for e in items:
        a = np.asarray(e, dtype=np.float64)
        lens.append(-1 if a.ndim == 0 else a.size)
        parts.append(a.reshape(-1))
    return np.concatenate([[len(items)], lens] + parts).astype(np.float64), isinstance(row, (tuple, list))


def _unflatten_row(v, was_seq):
    ne = int(v[0])
    lens = [int(x) for x in v[1:1 + ne]]
    pos, out = 1 + ne, []
    for ln in lens:
        if ln < 0:
            out.append(np.float64(v[pos])); pos += 1   # (np.float64, not float: CPython >= 3.12 sums floats compensated)
        else:
            out.append(np.array(v[pos:pos + ln])); pos += ln
    return tuple(out) if was_seq else out[0]


def decode_sharded(decode_blocks, blocks, group=None):
    """Parity-mode multi-GPU decode of one chunk: every rank has drawn ALL `blocks` from the
    same host RNG stream; rank r decodes blocks r, r + world, ... with `decode_blocks`, the per-block result rows
    are exchanged (one MAX all-reduce for the row width, one SUM all-reduce over disjoint supports = all-gather)
    and returned in draw order on every rank, so the reference's sequential stop rule (sparc_ldpc.py:1217-1245)
    replays identically everywhere.  Rows are exact float64 copies: results equal the one-rank run bit for bit."""
    rank, world_size = group_info(group)
    if world_size == 1:
        return decode_blocks(blocks)
    g = None if group is True else group
    nb = len(blocks)
    mine = list(range(rank, nb, world_size))
    rows = decode_blocks([blocks[j] for j in mine]) if mine else []
    flat = [_flatten_row(r) for r in rows]
    dev = "cuda" if dist.get_backend(g) == "nccl" else "cpu"
    meta = torch.tensor([max([len(f[0]) for f in flat], default=0), 1 if (flat and flat[0][1]) else 0], dtype=torch.int64, device=dev)
    dist.all_reduce(meta, op=dist.ReduceOp.MAX, group=g)
    W, was_seq = int(meta[0]), bool(int(meta[1]))
    buf = torch.zeros((nb, W), dtype=torch.float64, device=dev)
    for j, f in zip(mine, flat):
        buf[j, :len(f[0])] = torch.from_numpy(f[0]).to(dev)
    dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=g)
    host = buf.cpu().numpy()
    return [_unflatten_row(host[j], was_seq) for j in range(nb)]
