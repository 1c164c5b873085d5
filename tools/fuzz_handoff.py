"""Randomised check of the two handoff kernels with every lane busy (sp2bp_llr_kernel16, bp2sp_prior_kernel512)
against the one-warp-per-section kernels they replace (SB_HANDOFF_V1=1 selects those at call time): the outputs must
be bit-identical, including NaN / +-DBL_MAX classes, section lists, ragged counts and offsets."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sparc_ldpc_b200 import engine as E  # noqa: E402


def both(fn):
    os.environ.pop("SB_HANDOFF_V1", None)
    new = fn()
    os.environ["SB_HANDOFF_V1"] = "1"
    old = fn()
    os.environ.pop("SB_HANDOFF_V1", None)
    return new, old


def same(a, b):
    a, b = a.cpu().numpy(), b.cpu().numpy()
    return np.array_equal(a.view(np.uint64), b.view(np.uint64))


rs = np.random.RandomState(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
bad = 0
for case in range(120):
    M = int(2 ** rs.randint(6, 11))
    L = int(rs.choice([1, 3, 7, 8, 9, 15, 16, 17, 40]))
    B = int(rs.choice([1, 2, 5]))
    n = int(rs.randint(8, 4000))
    Pl = torch.from_numpy(rs.rand(L) + 0.1).cuda()
    beta = rs.rand(B, L * M) ** 8
    beta[rs.rand(B, L * M) < 0.3] = 0.0
    hot = rs.randint(0, M, (B, L))
    for b in range(B):                                   # some one-hot sections: the saturated / NaN classes
        for l in range(L):
            if rs.rand() < 0.3:
                beta[b, l * M:(l + 1) * M] = 0.0
                beta[b, l * M + hot[b, l]] = np.sqrt(n * float(Pl[l]))
    beta = torch.from_numpy(beta).cuda()
    first = int(rs.randint(0, L))
    cnt = L - first
    new, old = both(lambda: E.sp2bp_llr(beta, M, n, Pl, beta_first=first, first_sec=first, out_first=0, count=cnt, want_p=True))
    ok = same(new[0], old[0]) and same(new[1], old[1])
    # section lists (threshold flow): ragged per-codeword lists
    nsec = torch.from_numpy(rs.randint(0, L + 1, B).astype(np.int32)).cuda()
    secs = np.zeros((B, L), dtype=np.int32)
    for b in range(B):
        k = int(nsec[b])
        secs[b, :k] = np.sort(rs.choice(L, k, replace=False))
    secs = torch.from_numpy(secs).cuda()
    out_a = torch.zeros((B, L * int(np.log2(M))), dtype=torch.float64, device="cuda")
    out_b = out_a.clone()

    def run(o):
        return E.sp2bp_llr(beta, M, n, Pl, sections=secs, nsec=nsec, out=o)
    os.environ.pop("SB_HANDOFF_V1", None); run(out_a)
    os.environ["SB_HANDOFF_V1"] = "1"; run(out_b); os.environ.pop("SB_HANDOFF_V1", None)
    ok = ok and same(out_a, out_b)
    if M == 512:
        ls = int(rs.randint(0, L + 1))
        app = torch.from_numpy(rs.randn(B, max(ls, 1) * 9) * rs.choice([1.0, 30.0, 800.0])).cuda()
        prev = torch.from_numpy(rs.rand(B, L * M)).cuda()
        for sc in (True, False):
            new, old = both(lambda: E.bp2sp_prior(app, ls, prev, L, M, n, Pl, scale_by_power=sc))
            ok = ok and same(new, old)
        prob = torch.from_numpy(rs.rand(B, max(ls, 1) * 9)).cuda()
        new, old = both(lambda: E.bp2sp_prior(prob, ls, prev, L, M, n, Pl, scale_by_power=False, from_prob=True))
        ok = ok and same(new, old)
    if not ok:
        bad += 1
        print("MISMATCH case %d: M=%d L=%d B=%d n=%d first=%d" % (case, M, L, B, n, first))
torch.cuda.synchronize()
print("fuzz_handoff: %d mismatching cases of 120" % bad)
sys.exit(1 if bad else 0)
