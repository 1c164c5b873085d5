"""Times the two handoff kernels that stream beta at the headline shape (L = M = 512): sp2bp + LLR (one read of beta)
and bp2sp + prior scaling (one write of beta).  SB_HANDOFF_V1=1 selects the one-warp-per-section kernels."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sparc_ldpc_b200 import engine as E  # noqa: E402

B, L, M, n = int(os.environ.get("PH_B", 1184)), 512, 512, 4608
rs = np.random.RandomState(0)
Pl = torch.full((L,), 4.0 / L, dtype=torch.float64, device="cuda")
beta = torch.rand(B, L * M, dtype=torch.float64, device="cuda")
beta = beta / beta.view(B, L, M).sum(-1, keepdim=True).repeat_interleave(M, -1).view(B, L * M) * float(np.sqrt(n * 4.0 / L))
app = torch.randn(B, L * 9, dtype=torch.float64, device="cuda") * 8
gb = B * L * M * 8 / 1e9
for name, fn in (("sp2bp_llr", lambda: E.sp2bp_llr(beta, M, n, Pl, count=L)),
                 ("bp2sp_prior", lambda: E.bp2sp_prior(app, L, beta, L, M, n, Pl))):
    best = 1e9
    for rep in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    print("%-12s %7.3f ms for %d codewords: %.0f GB/s of beta traffic (%.1f GB)" % (name, best, B, gb / best * 1e3, gb))
