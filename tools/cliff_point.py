"""The bench's operating point on the cliff of the waterfall (7.667 ref-dB, L = M = 512, 802.16 5/6, soft exchange x2;
ldpc/EbN0_dBVsBER_waterfallsoft_rep200_LM512p4r1rldpc5_6.csv row 8), settled two ways:

 A. modes: the SAME device-generated codewords (default 4736) decoded in strict/strict, f64/strict and fast/fast;
    per-stage BER and block-failure counts of every mode, and the decisions of every codeword whose decodes all
    converge (AMP early stop in all three AMP calls, BP < 200 iterations in both BP calls, judged on the strict run)
    compared codeword by codeword between the modes -- they must be identical.
 B. oracle: K codewords (default 256) drawn from the reference's host RNG stream (codeword i from RandomState(seed0 + i),
    the reference's draw order) decoded by the CPU oracle (all host cores) and by the GPU in f64 and fast mode; the
    per-codeword (ber_amp[3], ber_ldpc[2]) tuples must be IDENTICAL except for codewords in the documented chaotic
    classes (a non-convergent AMP or BP decode in the flow), which are listed.

  python tools/cliff_point.py [--n 4736] [--k 256] [--json profiles/r02_cliff_point.json]

The oracle is the checker here (test infrastructure, like tools/fuzz_*.py): nothing in the product path imports it."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

L, M, P, T, SOFT = 512, 512, 4.0, 64, 2
EBN0 = 7.667
SIGMA = float(np.sqrt(P / (10 ** (EBN0 / 20) * 2 * (5.0 / 6.0))))
LDPC = ("802.16", "5/6", 192)
REF_ROW = {"EbN0_dB": 7.667, "BER_amp_1": 2.6e-3, "BER_ldpc": 9.8e-4, "blocks": "200-250",
           "source": "ldpc/EbN0_dBVsBER_waterfallsoft_rep200_LM512p4r1rldpc5_6.csv:8"}
MODES = (("strict", "strict"), ("f64", "strict"), ("fast", "fast"))


def _flow(su, y, tx, amp_mode, bp_mode, chunk=1184):
    from sparc_ldpc_b200 import decoder as D, engine as E
    prev = E.AMP_MODE, E.BP_MODE
    E.AMP_MODE, E.BP_MODE = amp_mode, bp_mode
    try:
        dec, errs, aexec, bit = [], [], [], []
        for c0 in range(0, y.shape[0], chunk):
            st = D.soft(su, y[c0:c0 + chunk], SOFT)
            stages = [st.amp_idx[0]]
            for j in range(SOFT):
                stages += [st.ldpc_idx[j], st.amp_idx[j + 1]]
            dec.append(torch.stack(stages).cpu().numpy())                                       # [5, b, L]
            errs.append(torch.stack([E.count_errors(s, tx[c0:c0 + chunk]) for s in stages]).cpu().numpy())
            aexec.append(torch.stack(st.amp_exec).cpu().numpy())                                # [3, b]
            bit.append(torch.stack(st.bp_it).cpu().numpy())                                     # [2, b]
    finally:
        E.AMP_MODE, E.BP_MODE = prev
    return (np.concatenate(dec, axis=1), np.concatenate(errs, axis=1), np.concatenate(aexec, axis=1),
            np.concatenate(bit, axis=1))


def part_a(n, seed=2024):
    from sparc_ldpc_b200 import decoder as D, montecarlo as MC, sparc_ldpc as S
    su = D.make_setup(S.SPARCParams(L=L, M=M, sigma=SIGMA, p=P, r=1, t=T), S.LDPCParams(*LDPC))
    gen = torch.Generator(device=su.dev)
    gen.manual_seed(seed)
    tx, y = MC.generate(su, n, SIGMA, gen)
    out, res = {}, {}
    for am, bm in MODES:
        t0 = time.time()
        dec, errs, aexec, bit = _flow(su, y, tx, am, bm)
        out[am] = (dec, errs, aexec, bit)
        res[am + "/" + bm] = {
            "ber_per_stage[amp1,ldpc1,amp2,ldpc2,amp3]": (errs.sum(axis=1) / (n * su.total_bits)).tolist(),
            "block_failures_per_stage": (errs > 0).sum(axis=1).tolist(),
            "mean_amp_iterations_per_decode": float(aexec.mean()), "mean_bp_iterations_per_decode": float(bit.mean()),
            "wall_s": time.time() - t0}
    s = out["strict"]
    conv = (s[2] < T).all(axis=0) & (s[3] < 200).all(axis=0)
    cmp_ = {}
    for am in ("f64", "fast"):
        o = out[am]
        same_cw = (o[0] == s[0]).all(axis=(0, 2))                    # decisions of all 5 stages equal
        cmp_[am + "_vs_strict"] = {
            "converged_codewords": int(conv.sum()), "converged_with_identical_decisions": int((same_cw & conv).sum()),
            "non_converged_codewords": int((~conv).sum()), "non_converged_with_identical_decisions": int((same_cw & ~conv).sum()),
            "block_failures_final_stage[strict,%s]" % am: [int((s[1][-1] > 0).sum()), int((o[1][-1] > 0).sum())],
            "block_failures_final_stage_among_converged[strict,%s]" % am: [int(((s[1][-1] > 0) & conv).sum()),
                                                                          int(((o[1][-1] > 0) & conv).sum())]}
    return {"codewords": n, "generator_seed": seed, "modes": res, "comparison": cmp_, "ok": part_a_ok(cmp_)}


def part_a_ok(cmp_):
    """f64 differs from strict by fp64 summation-order noise only: every converged codeword must decide identically.
    fast stops earlier (|d tau| <= 2^-27 tau) on 27-bit gathers: a converged codeword may flip a near-tie section
    (documented class; allowed: at most 0.2 % of the converged codewords, and never a change of the failure counts
    beyond those codewords)."""
    f, q = cmp_["f64_vs_strict"], cmp_["fast_vs_strict"]
    flips = q["converged_codewords"] - q["converged_with_identical_decisions"]
    q["near_tie_flips_among_converged"] = flips
    return bool(f["converged_codewords"] == f["converged_with_identical_decisions"]
                and flips <= max(1, 0.002 * q["converged_codewords"]))


def _oracle_one(seed):
    import warnings
    warnings.simplefilter("ignore")
    from oracle import oracle as orc
    rec = {}
    a, l_, _ = orc.soft_amp_ldpc_sim(orc.SPARCParams(L=L, M=M, sigma=SIGMA, p=P, r=1, t=T), orc.LDPCParams(*LDPC), SOFT,
                                     rng=np.random.RandomState(seed), record=rec)
    return seed, a, l_, [int(i) for i in rec["it"]]


def part_b(k, seed0=5000, cores=None):
    import multiprocessing as mp
    import subprocess
    from sparc_ldpc_b200 import decoder as D, sparc_ldpc as S
    if not os.path.isfile(os.path.join(ROOT, "oracle", "_build", "liboracle.so")):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "_build/liboracle.so"], stdout=subprocess.DEVNULL)
    cores = cores or os.cpu_count() or 1
    seeds = [seed0 + i for i in range(k)]
    t0 = time.time()
    with mp.get_context("fork").Pool(min(cores, k)) as pool:
        ref = pool.map(_oracle_one, seeds, chunksize=1)
    t_cpu = time.time() - t0
    su = D.make_setup(S.SPARCParams(L=L, M=M, sigma=SIGMA, p=P, r=1, t=T), S.LDPCParams(*LDPC))
    idx = np.empty((k, L), dtype=np.int32)
    noise = np.empty((k, su.n))
    for i, sd in enumerate(seeds):                                   # codeword i from its own stream, reference draw order
        idx[i], noise[i] = (v[0] for v in S._draw(su, 1, SIGMA, np.random.RandomState(sd)))
    tx, y = S._transmit(su, idx, noise)
    out = {}
    for am, bm in (("f64", "strict"), ("fast", "fast")):
        t0 = time.time()
        dec, errs, aexec, bit = _flow(su, y, tx, am, bm)
        ber = errs / su.total_bits                                   # [5, k]; the reference divides the same way (:650)
        rows, listed = 0, []
        for i, (sd, a, l_, it) in enumerate(ref):
            want = [a[0], l_[0], a[1], l_[1], a[2]]
            got = ber[:, i].tolist()
            if got == want:
                rows += 1
                continue
            chaotic = bool((aexec[:, i] >= T).any() or (bit[:, i] >= 200).any() or max(it) >= 200)
            listed.append({"seed": sd, "gpu": got, "oracle": want, "gpu_amp_iterations": aexec[:, i].tolist(),
                           "gpu_bp_iterations": bit[:, i].tolist(), "oracle_bp_iterations": it,
                           "class": "non-convergent AMP or BP decode in the flow (DESIGN.md section 3, classes 2 / 5)" if chaotic
                                    else "UNEXPLAINED"})
        out[am + "/" + bm] = {"identical_tuples": rows, "different": len(listed),
                              "unexplained": sum(1 for x in listed if x["class"] == "UNEXPLAINED"),
                              "mean_ber_gpu": ber.mean(axis=1).tolist(),
                              "mean_ber_oracle": np.mean([[a[0], l_[0], a[1], l_[1], a[2]] for _, a, l_, _ in ref], axis=0).tolist(),
                              "different_codewords": listed, "wall_s": time.time() - t0}
    ok = all(v["unexplained"] == 0 for v in out.values())
    return {"codewords": k, "seed0": seed0, "oracle_wall_s": t_cpu, "oracle_cores": min(cores, k), "modes": out, "ok": bool(ok)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=4736)
    ap.add_argument("--k", type=int, default=256)
    ap.add_argument("--json", default="")
    args = ap.parse_args()
    torch.cuda.set_device(0)
    rec = {"point": {"EbN0_ref_dB": EBN0, "sigma": SIGMA, "L": L, "M": M, "P": P, "ldpc": LDPC, "soft_iter": SOFT},
           "reference_csv_row": REF_ROW}
    if args.n:
        rec["A_same_codewords_three_modes"] = part_a(args.n)
    if args.k:
        rec["B_reference_stream_vs_oracle"] = part_b(args.k)
    s = json.dumps(rec, indent=1)
    if args.json:
        with open(args.json, "w") as fh:
            fh.write(s)
    brief = json.loads(s)
    for m in brief.get("B_reference_stream_vs_oracle", {}).get("modes", {}).values():
        m.pop("different_codewords", None)
    print(json.dumps(brief, indent=1))
    ok = all(rec[k]["ok"] for k in rec if k[:2] in ("A_", "B_"))
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
