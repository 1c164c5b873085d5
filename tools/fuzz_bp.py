"""Randomised check of the BP kernels over every protograph family in the code tables: random z, noisy channel LLRs
at a noise level where most blocks converge, GPU (sumprod2 strict and fast, minsum) against the CPU oracle
(restatement of c_ldpc.c).  Convergent blocks: same iteration count, same decisions, app to 1e-8 (strict)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import oracle as orc  # noqa: E402
from sparc_ldpc_b200 import ldpc  # noqa: E402

rs = np.random.RandomState(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
keys = list(ldpc._db().keys())
bad, n_blocks, n_conv = 0, 0, 0
for key in keys:
    std, rate, pt, _ = key.split("|")
    pt = "A" if pt == "-" else pt
    for z in (int(rs.choice([8, 12, 27, 33, 54])), int(rs.choice([81, 96, 100]))):
        try:
            c = ldpc.code(std, rate, z, pt)
            co = orc.Code(std, rate, z, pt)
        except Exception as ex:  # a (family, z) the reference rejects as well
            continue
        B = 6
        try:
            X = c.encode_batch(rs.randint(0, 2, (B, c.K)))
        except NameError:   # the custom protographs have no encoder; the reference sends the all-zero word (sparc_ldpc.py:930-941)
            X = np.zeros((B, c.N), dtype=np.int64)
        R = c.K / c.N
        s = float(np.sqrt(1.0 / (2 * R * 10 ** (rs.uniform(2.0, 4.5) / 10))))
        ch = 2 / s ** 2 * (1 - 2.0 * X + s * rs.randn(*X.shape))
        g = c.graph()
        a0, i0 = g.bp(torch.from_numpy(ch).cuda(), "sumprod2")
        a1, i1 = g.bp(torch.from_numpy(ch).cuda(), "sumprod2_fast")
        a0, i0, a1, i1 = a0.cpu().numpy(), i0.cpu().numpy(), a1.cpu().numpy(), i1.cpu().numpy()
        for b in range(B):
            ao, io = co.decode(ch[b])
            n_blocks += 1
            if io >= 200:
                continue
            n_conv += 1
            ok = io == i0[b] and np.array_equal(ao < 0, a0[b] < 0) and np.allclose(a0[b], ao, rtol=1e-8, atol=1e-8)
            okf = (i1[b] < 200) and np.array_equal(a1[b] < 0, ao < 0) and (io > 25 or i1[b] == io)
            if not (ok and okf):
                bad += 1
                print("MISMATCH %s z=%d block %d: it oracle %d strict %d fast %d, max app err %.2e"
                      % (key, z, b, io, i0[b], i1[b], np.max(np.abs(a0[b] - ao) / (1 + np.abs(ao)))))
print("fuzz_bp: %d code instances x 6 blocks = %d blocks, %d convergent, %d mismatches" % (n_blocks // 6, n_blocks, n_conv, bad))
sys.exit(1 if bad else 0)
