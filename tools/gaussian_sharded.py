"""Column-sharded Gaussian-mode AMP over NCCL: rank r holds the columns of sections [r L/W, (r+1) L/W) of a dense
design matrix; one all-reduce of [B*(n+1)] doubles per AMP iteration (sb_dense_amp_batch_sharded).

  torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/gaussian_sharded.py [--L 2048 --M 32 --n 10240 --B 128]

Every rank generates the same seeded matrix column block by block, so no rank ever holds the whole A."""
import argparse
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sparc_ldpc_b200 import engine as E  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--L", type=int, default=2048)
ap.add_argument("--M", type=int, default=32)
ap.add_argument("--n", type=int, default=10240)
ap.add_argument("--B", type=int, default=128)
ap.add_argument("--P", type=float, default=4.0)
ap.add_argument("--sigma", type=float, default=0.7)
ap.add_argument("--T", type=int, default=32)
args = ap.parse_args()

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
L, M, n, B = args.L, args.M, args.n, args.B
assert L % world == 0
Ll = L // world
Pl = torch.full((L,), args.P / L, dtype=torch.float64, device="cuda")
# the codewords: y = A beta0 + noise, built shard by shard with the same collective the decoder uses
gen = torch.Generator(device="cuda")
blocks = []
for r in range(world):
    gen.manual_seed(1234 + r)
    blk = torch.randn((n, Ll * M), dtype=torch.float64, device="cuda", generator=gen) / np.sqrt(n)
    if r == rank:
        A_local = blk
    del blk
gen.manual_seed(99)
idx = torch.randint(0, M, (B, L), device="cuda", generator=gen)
noise = torch.randn((B, n), dtype=torch.float64, device="cuda", generator=gen) * args.sigma
b0 = torch.zeros((B, Ll * M), dtype=torch.float64, device="cuda")
mine = idx[:, rank * Ll:(rank + 1) * Ll]
b0.scatter_(1, (torch.arange(Ll, device="cuda") * M)[None, :] + mine, float(np.sqrt(n * args.P / L)))
x = b0 @ A_local.t()
if world > 1:
    dist.all_reduce(x)
y = x + noise
op = E.DenseOperator(A_local, Ll, M)
for rep in range(3):
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    if world > 1:
        res = op.amp_sharded(y, Pl[rank * Ll:(rank + 1) * Ll].contiguous(), args.P, args.T)
    else:
        res = op.amp(y, Pl, args.T)
    e1.record()
    torch.cuda.synchronize()
    dec = res.beta.view(B, Ll, M).argmax(dim=2)
    errs = (dec != mine).sum().to(torch.float64)
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(errs)
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        its = int(res.n_exec.max())
        print("world %d rep %d: %.2f ms for %d codewords x %d iterations (L=%d M=%d n=%d, %d sections per rank), "
              "section error rate %.4f, all-reduce of %.2f MB per iteration"
              % (world, rep, float(ms), B, its, L, M, n, Ll, float(errs) / (B * L), (B * n + B) * 8 / 1e6))
if world > 1:
    dist.destroy_process_group()
