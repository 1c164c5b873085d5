"""Profiling / timing driver of the Gaussian-mode tcgen05 GEMM (csrc/dense.cu) at a large shape.
python tools/profile_dense.py [--n 4608 --LM 65536 --B 256 --reps 5]"""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sparc_ldpc_b200 import engine as E  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=4608)
ap.add_argument("--LM", type=int, default=65536)
ap.add_argument("--B", type=int, default=256)
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--check", action="store_true")
args = ap.parse_args()

torch.manual_seed(0)
A = torch.randn((args.n, args.LM), dtype=torch.float64, device="cuda") / np.sqrt(args.n)
op = E.DenseOperator(A, 1, args.LM)
x = torch.randn((args.B, args.LM), dtype=torch.float64, device="cuda")
z = torch.randn((args.B, args.n), dtype=torch.float64, device="cuda")
if args.check:
    e1 = (op.Ab(x) - x @ A.t()).abs().max() / (x.abs() @ A.t().abs()).max()
    e2 = (op.Az(z) - z @ A).abs().max() / (z.abs() @ A.abs()).max()
    print("max err / max(|A||x|): A x %.2e, A^T z %.2e" % (float(e1), float(e2)))
flop = 2.0 * args.n * args.LM * args.B
for name, f, v in (("A x", op.Ab, x), ("A^T z", op.Az, z)):
    for i in range(args.reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        f(v)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        print("%s rep %d: %.3f ms (split + GEMM + combine), %.1f TFLOP/s useful fp32-equivalent, %.1f TFLOP/s bf16 MMA issued"
              % (name, i, ms, flop / ms / 1e9, 6 * flop / ms / 1e9))
