"""Profiling driver: a few launches of the AMP kernel (and optionally the BP kernel) at the headline shape.
Used under ncu on the GPU box:  ncu --set full -k regex:amp_kernel -s 2 -c 1 python tools/profile_amp.py"""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sparc_ldpc_b200 import decoder as D, sparc_ldpc as S  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=296)
ap.add_argument("--T", type=int, default=8)
ap.add_argument("--launches", type=int, default=3)
ap.add_argument("--sigma", type=float, default=0.9964)
ap.add_argument("--bp", action="store_true")
ap.add_argument("--mode", default="fast")
ap.add_argument("--r", type=float, default=1.0, help="SPARC rate (n = L logM / r)")
ap.add_argument("--bp-rule", default="sumprod2")
args = ap.parse_args()

sp = S.SPARCParams(L=512, M=512, sigma=args.sigma, p=4.0, r=args.r, t=args.T)
su = D.make_setup(sp, S.LDPCParams("802.16", "5/6", 192))
idx, noise = S._draw(su, args.batch, args.sigma, np.random.RandomState(0))
tx, y = S._transmit(su, idx, noise)
torch.cuda.synchronize()
for i in range(args.launches):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    res = su.op.amp(y, su.Pl_dev, args.T, mode=args.mode)
    e1.record()
    torch.cuda.synchronize()
    it = float(res.n_exec.sum())
    ms = e0.elapsed_time(e1)
    print("amp launch %d: %.2f ms, %d codeword-iterations, %.2f us per codeword-iteration, %.1f GB/s algorithmic"
          % (i, ms, it, 1e3 * ms / it, it * (2 * 512 * 512 + 3 * su.n) * 8 / ms / 1e6))
if args.bp:
    from sparc_ldpc_b200 import engine as E
    llr = E.sp2bp_llr(res.beta, 512, su.n, su.Pl_dev, count=512)
    for i in range(2):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        app, itb = su.graph.bp(llr, args.bp_rule)
        e1.record()
        torch.cuda.synchronize()
        print("bp launch %d: %.2f ms, %d iterations total" % (i, e0.elapsed_time(e1), int(itb.sum())))
