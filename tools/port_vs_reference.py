"""cpu_baseline.kind: "port" vs the unmodified reference, on record.  Runs ONE soft-flow codeword of the
bench workload (L = M = 512, 802.16 5/6 z = 192, 2 AMP<->BP rounds, 7.667 ref-dB, seed 1234) through
  (a) the UNMODIFIED reference imported by oracle/ref_harness.py (needs /root/reference; its w-point transform runs in the
      compiled transcription of sparc_ldpc.py:19-29 that the harness installs -- the reference's own pyfht extension is a
      compiled module too), and
  (b) the oracle port bench.py times as `cpu_baseline` / `--impl reference`,
on one core each, checks that both return the same BER tuple, and writes the times.  The reference tree does not exist on
the GPU box, so this runs in the build container:  python tools/port_vs_reference.py [--json profiles/r02_port_vs_reference.json]"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
ap = argparse.ArgumentParser()
ap.add_argument("--json", default="")
args = ap.parse_args()
out = os.path.abspath(args.json) if args.json else ""

L, M, P, T = 512, 512, 4.0, 64
SIGMA = float(np.sqrt(P / (10 ** (7.667 / 20) * 2 * (5.0 / 6.0))))
from oracle import oracle as orc            # noqa: E402
from oracle import ref_harness              # noqa: E402

t0 = time.perf_counter()
port = orc.soft_amp_ldpc_sim(orc.SPARCParams(L=L, M=M, sigma=SIGMA, p=P, r=1, t=T), orc.LDPCParams("802.16", "5/6", 192), 2,
                             rng=np.random.RandomState(1234))
t_port = time.perf_counter() - t0
rec = {"workload": "one soft-flow codeword, L=M=512 R=1 P=4, 802.16 5/6 z=192, 2 AMP<->BP rounds, Eb/N0(ref dB)=7.667, seed 1234",
       "port_s": t_port, "port_result": [port[0], port[1]], "cores": 1}
if ref_harness.reference_available():
    sl, ae, at, ldpc = ref_harness.load_reference()
    np.random.seed(1234)
    t0 = time.perf_counter()
    ref = sl.soft_amp_ldpc_sim(sl.SPARCParams(L, M, SIGMA, P, 1, T), sl.LDPCParams("802.16", "5/6", 192), 2)
    t_ref = time.perf_counter() - t0
    rec.update(reference_s=t_ref, reference_result=[list(map(float, ref[0])), list(map(float, ref[1]))],
               identical_ber_tuples=bool(list(map(float, ref[0])) == port[0] and list(map(float, ref[1])) == port[1]),
               reference_over_port=t_ref / t_port)
else:
    rec["reference_s"] = None
print(json.dumps(rec, indent=1))
if out:
    with open(out, "w") as fh:
        json.dump(rec, fh, indent=1)
