#!/bin/bash
# compute-sanitizer memcheck over small decodes of every AMP kernel family and the handoff / BP kernels
mkdir -p gpurun_out
cat > /tmp/san.py <<'P'
import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from sparc_ldpc_b200 import decoder as D, engine as E, sparc_ldpc as S, montecarlo as MC
for (L, M, r, lp, B) in ((16, 512, 1, None, 5), (24, 512, 1, ("802.16", "5/6", 9), 3), (32, 32, 1, None, 5), (21, 8, 1, None, 3)):
    su = D.make_setup(S.SPARCParams(L=L, M=M, sigma=0.9, p=4.0, r=r, t=6), None if lp is None else S.LDPCParams(*lp))
    gen = torch.Generator(device=su.dev); gen.manual_seed(1)
    tx, y = MC.generate(su, B, 0.9, gen)
    for mode in ("strict", "f64", "fast"):
        E.AMP_MODE = mode
        st = (D.soft(su, y, 1) if lp else D.plain(su, y))
        torch.cuda.synchronize()
    print("ok", L, M, lp)
P
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 7 python /tmp/san.py > gpurun_out/r2_sanitizer.log 2>&1; echo "sanitizer rc=$?" | tee -a gpurun_out/r2_sanitizer.log
tail -12 gpurun_out/r2_sanitizer.log
