"""BER waterfall of the GPU decoder next to the curves the reference published (tests/golden/reference_waterfall.json
= ldpc/EbN0_dBVsBER_waterfall{soft,hard,OriginalHard}_rep200_LM512p4r1rldpc5_6.csv of Spimp/sparc_ldpc).
L=M=512, P=4, r=1, 802.16 rate-5/6 (z=192); codewords are generated on the device (throughput mode), the Eb/N0 grid
and the sigma convention are the reference's (sparc_ldpc.py:1158-1160,1169,1184,1199-1200).

  python tools/waterfall_vs_reference.py [--n 2368] [--flow soft] [--out profiles/r01_waterfall_soft.json]
  torchrun --nproc-per-node 8 tools/waterfall_vs_reference.py --n 10000       # sharded over the GPUs"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sparc_ldpc_b200 import engine as E, montecarlo as MC, sparc_ldpc as S  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=2368, help="codewords per Eb/N0 point (all ranks together)")
ap.add_argument("--flow", default="soft", choices=["soft", "hard", "originalHard"])
ap.add_argument("--amp-mode", default="fast")
ap.add_argument("--bp-mode", default="fast", help="engine.BP_MODE: strict | fast (the bench runs fast)")
ap.add_argument("--out", default="")
args = ap.parse_args()
E.BP_MODE = args.bp_mode
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    import torch.distributed as dist
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ref = json.load(open(os.path.join(ROOT, "tests", "golden", "reference_waterfall.json")))[args.flow]
L, M, P, R = 512, 512, 4.0, 5.0 / 6.0
lp = S.LDPCParams("802.16", "5/6", 192)
rows = []
t0 = time.time()
for i, r in enumerate(ref["rows"]):
    db = r["EbN0_dB"]
    sigma = float(np.sqrt(P / (10 ** (db / 20) * 2 * R)))
    sp = S.SPARCParams(L=L, M=M, sigma=sigma, p=P, r=1, t=64)
    kw = dict(soft_iter=2) if args.flow == "soft" else {}
    res = MC.ber_point(sp, lp, args.n, flow=args.flow, seed=100 + i, amp_mode=args.amp_mode, **kw)
    # plain SPARC at the same overall rate (r = 5/6 -> n = 5529), as the reference's waterfall() runs it (:1231)
    plain = MC.ber_point(S.SPARCParams(L=L, M=M, sigma=sigma, p=P, r=R, t=64), None, args.n, flow="plain", seed=500 + i,
                         amp_mode=args.amp_mode)
    row = {"EbN0_dB": db, "sigma": sigma, "n_codewords": res["n_codewords"],
           "ours": {"BER_amp": res["ber_amp"], "BER_ldpc": res["ber_ldpc"], "BER_plain": plain["ber_amp"][0],
                    "block_errors_ldpc": res["block_errors_ldpc"]},
           "reference": {k: v for k, v in r.items() if k != "EbN0_dB"}}
    rows.append(row)
    if rank == 0:
        print("%.3f dB  ours amp %s ldpc %s plain %.3e | reference amp1 %.3e ldpc %.3e amp2 %.3e ldpc2 %.3e plain %.3e"
              % (db, ["%.3e" % v for v in res["ber_amp"]], ["%.3e" % v for v in res["ber_ldpc"]], plain["ber_amp"][0],
                 r["BER_amp_1"], r["BER_ldpc"], r["BER_amp_2"], r["BER_ldpc_2"], r["BER_plain"]), flush=True)
if rank == 0:
    out = {"flow": args.flow, "reference_csv": ref["source"], "codewords_per_point": args.n, "n_gpus": world,
           "amp_mode": args.amp_mode, "bp_mode": args.bp_mode, "wall_s": time.time() - t0, "rows": rows}
    if args.out:
        json.dump(out, open(args.out, "w"), indent=1)
    print("wall %.1f s for %d points x %d codewords (coded + plain) on %d GPU(s)" % (out["wall_s"], len(rows), args.n, world))
if world > 1:
    dist.destroy_process_group()
