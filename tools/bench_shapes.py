"""Throughput of every BASELINE.json shape (north star: "throughput on synthetic codewords of each named
(L, M, R, P) shape"), one JSON line per shape.  Codewords are generated on the device (montecarlo.generate), the
timed region is the decode flow only (CUDA events, after one warm-up batch); `frac` is the AMP kernel's
algorithmic HBM traffic ((2 L M + 3 n) * 8 bytes per executed codeword-iteration) over the whole flow's time,
against MEASURED_PEAKS.json.  Under torchrun every rank decodes its own batch (weak scaling) and the counters are
all-reduced.

  python tools/bench_shapes.py [--batch 1184] [--reps 2] [--only C1,C3]"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sparc_ldpc_b200 import decoder as D, engine as E, montecarlo as MC, sparc_ldpc as S  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=1184)
ap.add_argument("--reps", type=int, default=2)
ap.add_argument("--only", default="")
ap.add_argument("--amp-mode", default="fast")
ap.add_argument("--bp-mode", default="fast")
args = ap.parse_args()

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    import torch.distributed as dist
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
E.AMP_MODE = args.amp_mode
E.BP_MODE = args.bp_mode
try:
    PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    PEAK = 6650.0


def sigma_of(db, P, R):
    return float(np.sqrt(P / (10 ** (db / 20) * 2 * R)))       # sparc_ldpc.py:1184,1199-1200


# name -> (L, M, r, P, ldpc (standard, rate, z) | None, flow, kwargs, Eb/N0 in the reference's dB, note)
SHAPES = {
    "C1": (128, 4, 1, 2.0, None, "plain", {}, 6.0, "configs[0] plain SPARC AMP (structured operator)"),
    "C2": (512, 512, 1, 4.0, None, "plain", {}, 8.0, "configs[1] plain SPARC AMP, waterfall point"),
    "C3": (512, 512, 1, 4.0, ("802.16", "5/6", 192), "soft", {"soft_iter": 2}, 7.667, "configs[2] soft exchange x2 (= bench.py)"),
    "C3-hard": (512, 512, 1, 4.0, ("802.16", "5/6", 192), "hard", {}, 7.667, "configs[2] hard-beta init"),
    "C3-threshold": (512, 512, 1, 4.0, ("802.16", "5/6", 192), "threshold", {"soft_iter": 2, "thr": 0.6}, 9.43, "configs[2] threshold init 0.6"),
    "C4": (256, 32, 1, 4.0, None, "plain", {}, 11.0, "configs[3] shape (EXIT chart runs AMP on peeled section lists of this operator)"),
    "C5": (768, 512, 5 / 6, 1.8, ("802.16", "1/2", 33), "soft", {"soft_iter": 2}, 7.27,
           "configs[4] with z = 33 (z = 32 violates the reference's own precondition nl % logM == 0, SURVEY 8d)"),
}
only = [s for s in args.only.split(",") if s]
for name, (L, M, r, P, lp, flow, kw, db, note) in SHAPES.items():
    if only and name not in only:
        continue
    lpp = None if lp is None else S.LDPCParams(*lp)
    su = D.make_setup(S.SPARCParams(L=L, M=M, sigma=1.0, p=P, r=r, t=64), lpp)
    sigma = sigma_of(db, P, su.R)
    gen = torch.Generator(device=su.dev)
    gen.manual_seed(7 + rank)
    B = args.batch
    tx, y = MC.generate(su, B, sigma, gen)
    f = MC.FLOWS[flow]
    f(su, y, **kw)                                              # warm-up (also builds / caches the tables)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.reps):
        st = f(su, y, **kw)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.reps
    its = float(sum(int(a.sum()) for a in st.amp_exec))
    errs = E.count_errors(st.ldpc_idx[-1] if st.ldpc_idx else st.amp_idx[-1], tx).sum().to(torch.float64)
    t = torch.tensor([ms, its, float(errs)], dtype=torch.float64, device=su.dev)
    if world > 1:
        mx = t.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        dist.all_reduce(t)
        ms, its, errs = float(mx[0]), float(t[1]), float(t[2])
    else:
        ms, its, errs = float(t[0]), float(t[1]), float(t[2])
    if rank == 0:
        bytes_it = (2 * L * M + 3 * su.n) * 8
        info = su.total_bits - (su.nl - su.kl)
        print(json.dumps({
            "shape": name, "note": note, "L": L, "M": M, "n": su.n, "P": P, "flow": flow, "ldpc": lp, "EbN0_ref_dB": db,
            "sigma": sigma, "n_gpus": world, "codewords_per_gpu": B, "ms": ms, "codewords_per_s": B * world / (ms / 1e3),
            "info_mbit_per_s": B * world * info / (ms / 1e3) / 1e6,
            "amp_iterations_per_codeword": its / (B * world), "us_per_codeword_iteration": 1e3 * ms * world / max(its, 1),
            "algorithmic_GBps_per_gpu": its / world * bytes_it / (ms / 1e3) / 1e9,
            "frac_of_measured_hbm": its / world * bytes_it / (ms / 1e3) / 1e9 / PEAK,
            "final_ber": errs / (B * world * su.total_bits), "amp_mode": args.amp_mode, "bp_mode": args.bp_mode}))
if world > 1:
    dist.destroy_process_group()
