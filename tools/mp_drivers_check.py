"""Multi-rank check of the parity-mode Monte-Carlo drivers:

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
      tools/mp_drivers_check.py [--same-gpu] [--backend gloo|nccl]

Every rank first runs each driver alone (group=None), then all ranks run it together (group=True): the CSV columns and
the final state of the host RNG stream must be IDENTICAL -- blocks are decoded by rank j mod world, rows are
exchanged exactly and the reference's sequential stop rule is replayed on every rank (ldpc/sparc_ldpc.py:1217-1251,
ldpc/amp_exit.py:560-595).  --same-gpu: all ranks share GPU 0 (gloo for the exchange); default: rank r uses GPU r.
Prints one JSON line per driver on rank 0 and exits non-zero on any mismatch."""
import argparse
import json
import os
import sys
import tempfile
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

ap = argparse.ArgumentParser()
ap.add_argument("--same-gpu", action="store_true")
ap.add_argument("--backend", default=None)
ap.add_argument("--big", action="store_true", help="also time a BASELINE configs[2]-sized waterfall point")
args = ap.parse_args()
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
local = 0 if args.same_gpu else int(os.environ.get("LOCAL_RANK", rank))
torch.cuda.set_device(local)
backend = args.backend or ("gloo" if args.same_gpu else "nccl")
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
if backend == "nccl":
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
else:
    dist.init_process_group("gloo")

from sparc_ldpc_b200 import amp_exit as AE, sparc_ldpc as S  # noqa: E402

tmp = tempfile.mkdtemp()


def same_state(a, b):
    return a[0] == b[0] and np.array_equal(a[1], b[1]) and a[2:] == b[2:]


def flat(d):
    return np.concatenate([np.asarray(d[k], dtype=float).reshape(-1) for k in sorted(d)])


def run_pair(name, fn):
    r1, r2 = np.random.RandomState(11), np.random.RandomState(11)
    t0 = time.time()
    single = fn(r1, None, os.path.join(tmp, "%s_r%d_single.csv" % (name, rank)))
    t1 = time.time()
    dist.barrier()
    t2 = time.time()
    shard = fn(r2, True, os.path.join(tmp, "%s_shared.csv" % name))
    dist.barrier()
    t3 = time.time()
    ok = np.array_equal(flat(single), flat(shard), equal_nan=True) and same_state(r1.get_state(), r2.get_state())
    if not ok:
        print("rank %d %s MISMATCH: rng state equal %s" % (rank, name, same_state(r1.get_state(), r2.get_state())), flush=True)
        for k in sorted(single):
            a, b = np.asarray(single[k], dtype=float), np.asarray(shard[k], dtype=float)
            if not np.array_equal(a, b, equal_nan=True):
                print("   %s single %s sharded %s" % (k, a.reshape(-1)[:8], b.reshape(-1)[:8]), flush=True)
    flag = torch.tensor([1 if ok else 0])
    if backend == "nccl":
        flag = flag.cuda()
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        rows = None
        if os.path.isfile(os.path.join(tmp, "%s_shared.csv" % name)):
            rows = sum(1 for _ in open(os.path.join(tmp, "%s_shared.csv" % name)))
        print(json.dumps({"driver": name, "world": world, "backend": backend, "identical_on_all_ranks": bool(int(flag)),
                          "single_rank_s": round(t1 - t0, 3), "sharded_s": round(t3 - t2, 3), "csv_lines_written_by_rank0": rows}),
              flush=True)
    return bool(int(flag))


lp = S.LDPCParams("802.16", "5/6", None)
sp = S.SPARCParams(64, 32, None, 4.0, 1, 32)
ok = True
ok &= run_pair("waterfall_soft", lambda r, g, f: S.waterfall(sp, lp, f, None, init="soft", MIN_ERRORS=5, MAX_BLOCKS=23,
                                                               sections=48, chunk=7, EbN0_dB=[6.0, 8.5], rng=r, group=g))
ok &= run_pair("waterfall_originalHard", lambda r, g, f: S.waterfall(sp, lp, f, None, init="originalHard", MIN_ERRORS=4,
                                                                       MAX_BLOCKS=17, sections=48, chunk=5, EbN0_dB=[7.0],
                                                                       bpsk=False, rng=r, group=g))
ok &= run_pair("soft_hard_plot", lambda r, g, f: S.soft_hard_plot(True, True, 48, 2, sp, lp, f, None, MIN_ERRORS=4,
                                                                    MAX_BLOCKS=13, chunk=5, SIGMA=[0.9], rng=r, group=g))
ok &= run_pair("soft_hardinit_plot", lambda r, g, f: S.soft_hardinit_plot(sp, lp, f, None, 48, MIN_ERRORS=4, MAX_BLOCKS=11,
                                                                            soft_iter=2, threshold=0.6, chunk=4, SIGMA=[1.0],
                                                                            rng=r, group=g))
ok &= run_pair("sim_ldpc", lambda r, g, f: {"ber": S.sim_ldpc(S.LDPCParams("802.16", "5/6", 8), 0.55, MIN_ERRORS=6,
                                                                MAX_BLOCKS=300, chunk=37, rng=r, group=g)})


def exit_curve(r, g, f):
    Ia, Ie, poly = AE.amp_exit_curve(S.SPARCParams(64, 8, None, 4.0, 1, 64), 10, 13, 2, 3, 0.7, bin_number=40, chunk=7, rng=r,
                                     group=g)
    return {"Ia": Ia, "Ie": Ie, "poly": poly}


ok &= run_pair("amp_exit_curve", exit_curve)
if args.big:
    spb = S.SPARCParams(512, 512, None, 4.0, 1, 64)
    ok &= run_pair("waterfall_soft_C3_point", lambda r, g, f: S.waterfall(spb, lp, f, None, init="soft", MIN_ERRORS=8,
                                                                            MAX_BLOCKS=8 * world, bpsk=False, chunk=8 * world,
                                                                            EbN0_dB=[7.667], rng=r, group=g))
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
