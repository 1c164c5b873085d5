"""Experiment builds only (make EXTRA=-DSB_PHASE_CLOCKS): where do the cycles of the AMP kernel go?
Thread 0 of every CTA accumulates the cycles between marks (amp_impl.cuh SB_CLK); prints the share per phase."""
import ctypes as ct
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sparc_ldpc_b200 import _lib, decoder as D, sparc_ldpc as S  # noqa: E402

B, T = int(os.environ.get("PC_B", 296)), int(os.environ.get("PC_T", 8))
sp = S.SPARCParams(L=512, M=512, sigma=0.9964, p=4.0, r=1, t=T)
su = D.make_setup(sp, S.LDPCParams("802.16", "5/6", 192))
idx, noise = S._draw(su, B, 0.9964, np.random.RandomState(0))
tx, y = S._transmit(su, idx, noise)
lib = ct.CDLL(_lib.LIB_PATH)  # SPARC_B200_LIB selects the experiment build
out = (ct.c_ulonglong * 16)()
names = ["other", "fold", "fht1", "softmax+store", "fht2+F", "wait1", "gather", "wait2", "tau+quant", "zupdate"]
for rep in range(3):
    lib.sb_phase_cycles_read(out)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    res = su.op.amp(y, su.Pl_dev, T, mode="fast")
    e1.record()
    torch.cuda.synchronize()
    lib.sb_phase_cycles_read(out)
    it = float(res.n_exec.sum())
    v = np.array(list(out)[:10], dtype=np.float64)
    print("launch %d: %.2f ms, %.2f us per codeword-iteration; cycles per section (thread 0 of each CTA):"
          % (rep, e0.elapsed_time(e1), 1e3 * e0.elapsed_time(e1) / it))
    per = v / (it * 512)
    print("   " + ", ".join("%s %.0f" % (n, p) for n, p in zip(names, per)) + "  | total %.0f" % per.sum())
