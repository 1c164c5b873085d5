"""Experiment builds only (tools/ab_build2.sh tr -DP2_TRACE): per-pass timeline of the pair kernel."""
import ctypes as ct
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sparc_ldpc_b200 import _lib, decoder as D, sparc_ldpc as S  # noqa: E402

B, T = int(os.environ.get("PC_B", 1184)), int(os.environ.get("PC_T", 64))
sp = S.SPARCParams(L=512, M=512, sigma=0.9964, p=4.0, r=1, t=T)
su = D.make_setup(sp, S.LDPCParams("802.16", "5/6", 192))
idx, noise = S._draw(su, B, 0.9964, np.random.RandomState(0))
tx, y = S._transmit(su, idx, noise)
lib = ct.CDLL(_lib.LIB_PATH)
lib.sb_p2_trace_read.argtypes = [ct.c_void_p, ct.c_int]
lib.sb_p2_trace_read.restype = ct.c_int
buf = np.zeros((1 << 20, 4), dtype=np.uint64)
for rep in range(2):
    lib.sb_p2_trace_read(buf.ctypes.data, 1 << 20)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    res = su.op.amp(y, su.Pl_dev, T, mode="fast")
    e1.record()
    torch.cuda.synchronize()
    n = lib.sb_p2_trace_read(buf.ctypes.data, 1 << 20)
    tr = buf[:n].astype(np.int64)
    np.save(os.path.join('gpurun_out', 'p2_trace_B%d_T%d_%d.npy' % (B, T, rep)), tr)
    t0 = tr[:, 0].min()
    start, endp, endb, meta = tr[:, 0] - t0, tr[:, 1] - t0, tr[:, 2] - t0, tr[:, 3]
    m0, m1 = ((meta >> 28) & 15) - 1, ((meta >> 24) & 15) - 1
    it0, it1 = (meta >> 8) & 0xFFFF, meta & 0xFF
    both = (m0 >= 0) & (m1 >= 0)
    print("launch %d: %.2f ms, %d passes (%d with both slots), codeword-iterations %d" % (rep, e0.elapsed_time(e1), n, both.sum(), int(res.n_exec.sum())))
    dp, db = (endp - start) / 1e3, (endb - endp) / 1e3
    print("  pass us: both mean %.1f (p10 %.1f p50 %.1f p90 %.1f) | single mean %.1f | boundary us: mean %.1f p50 %.1f p90 %.1f max %.1f"
          % (dp[both].mean(), *np.percentile(dp[both], [10, 50, 90]), dp[~both].mean() if (~both).any() else 0, db.mean(), *np.percentile(db, [50, 90]), db.max()))
    # pass duration vs iteration index of slot 0
    for lo, hi in ((0, 1), (1, 4), (4, 8), (8, 16), (16, 32), (32, 64)):
        sel = both & (it0 >= lo) & (it0 < hi)
        if sel.any():
            print("    slot-0 iteration %2d..%2d: %6d passes, mean %.1f us" % (lo, hi - 1, sel.sum(), dp[sel].mean()))
    print("  last pass ends at %.2f ms; sum(pass+boundary) per CTA: mean %.2f ms max %.2f ms" % (endb.max() / 1e6,
          np.bincount((meta >> 32).astype(np.int64), weights=(endb - start) / 1e6).mean(), np.bincount((meta >> 32).astype(np.int64), weights=(endb - start) / 1e6).max()))
