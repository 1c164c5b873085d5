"""Throughput of every BASELINE.json shape (north star: "throughput on synthetic codewords of each named
(L, M, R, P) shape"), one record per shape.  Codewords are generated on the device (montecarlo.generate), the
timed region is the decode flow only (CUDA events, after one warm-up batch); `frac` is the AMP kernel's
algorithmic HBM traffic ((2 L M + 3 n) * 8 bytes per executed codeword-iteration) over the whole flow's time,
against MEASURED_PEAKS.json.

Scaling per shape under torchrun: C1..C4 decode `batch` codewords on EVERY rank (weak scaling); C5 is BASELINE
configs[4], "10k codewords per Eb/N0 point sharded across 8 GPUs": the 10 000 codewords of ONE point are split
over the ranks (strong scaling), the error counters are all-reduced, and the time is the max over ranks.  C4 is
additionally run as one EXIT-chart point (calc_E on peeled section lists + histograms + I_e, amp_exit.py:185-351)
with the histogram all-reduce across ranks.

  python tools/bench_shapes.py [--batch 1184] [--reps 2] [--only C1,C3]          (one JSON line per shape)
  bench.py imports run_all() for the `shapes` block of its JSON line."""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def sigma_of(db, P, R):
    return float(np.sqrt(P / (10 ** (db / 20) * 2 * R)))       # sparc_ldpc.py:1184,1199-1200


def pa_exponential(L, P, R, db_first):
    """waterfall(pa_param=True): a = f = r / C with C = 0.5 log2(1 + P / sigma^2) of the FIRST grid point, frozen
    (sparc_ldpc.py:1201-1204); grid linspace(5.2, 10, 10) (the default grid crashes the reference)."""
    s0 = sigma_of(db_first, P, R)
    C = 0.5 * np.log2(1.0 + P / s0 ** 2)
    return dict(a=R / C, f=R / C, C=C)


# name -> (L, M, r, P, ldpc (standard, rate, z) | None, flow, kwargs, Eb/N0 in the reference's dB, note)
SHAPES = {
    "C1": (128, 4, 1, 2.0, None, "plain", {}, 6.0, "configs[0] plain SPARC AMP (structured operator)"),
    "C2": (512, 512, 1, 4.0, None, "plain", {}, 8.0,
           "configs[1] plain SPARC AMP with exponential power allocation (pa_param=True on linspace(5.2, 10, 10): a = f = r/C "
           "frozen at the first point), waterfall point 8.0 dB"),
    "C3": (512, 512, 1, 4.0, ("802.16", "5/6", 192), "soft", {"soft_iter": 2}, 7.667, "configs[2] soft exchange x2 (= bench.py)"),
    "C3-hard": (512, 512, 1, 4.0, ("802.16", "5/6", 192), "hard", {}, 7.667, "configs[2] hard-beta init"),
    "C3-threshold": (512, 512, 1, 4.0, ("802.16", "5/6", 192), "threshold", {"soft_iter": 2, "thr": 0.6}, 9.43, "configs[2] threshold init 0.6"),
    "C4": (256, 32, 1, 4.0, None, "plain", {}, 11.0, "configs[3] shape, plain AMP decode"),
    "C5": (768, 512, 5 / 6, 1.8, ("802.16", "1/2", 33), "soft", {"soft_iter": 2}, 7.27,
           "configs[4] with z = 33 (z = 32 violates the reference's own precondition nl % logM == 0, SURVEY 8d)"),
}
QUICK = ("C1", "C2", "C4", "C5")        # what bench.py's `shapes` block carries (C3 is the bench itself)
# codewords per rank = batch x this: enough waves of CTAs that the stragglers running all 64 iterations (C2) and the launch
# overhead of C4's 7 ms launches do not set the time
BATCH_FACTOR = {"C2": 4, "C4": 8}


def _reduce(vals, world, op="sum"):
    t = torch.tensor(vals, dtype=torch.float64, device="cuda")
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(t, op=dist.ReduceOp.MAX if op == "max" else dist.ReduceOp.SUM)
    return [float(v) for v in t]


def run_shape(name, rank, world, peak, batch, reps, amp_mode, bp_mode):
    from sparc_ldpc_b200 import decoder as D, engine as E, montecarlo as MC, sparc_ldpc as S
    E.AMP_MODE, E.BP_MODE = amp_mode, bp_mode
    L, M, r, P, lp, flow, kw, db, note = SHAPES[name]
    lpp = None if lp is None else S.LDPCParams(*lp)
    pa = pa_exponential(L, P, r, 5.2) if name == "C2" else {}
    su = D.make_setup(S.SPARCParams(L=L, M=M, sigma=1.0, p=P, r=r, t=64, **pa), lpp)
    sigma = sigma_of(db, P, su.R)
    gen = torch.Generator(device=su.dev)
    gen.manual_seed(7 + rank)
    scaling = "weak"
    B = batch * BATCH_FACTOR.get(name, 1)
    total = B * world
    if name == "C5":                    # 10 000 codewords of one Eb/N0 point, split over the ranks
        scaling, total = "strong", 10000
        B = total // world + (1 if rank < total % world else 0)
    tx, y = MC.generate(su, B, sigma, gen)
    f = MC.FLOWS[flow]
    f(su, y[: min(B, 296)], **kw)                               # warm-up (also builds / caches the tables)
    CH = 2500                                                   # codewords per launch chain: bounds the beta working set
    torch.cuda.synchronize()
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        errs_dev = torch.zeros((), dtype=torch.float64, device=su.dev)
        its_dev = torch.zeros((), dtype=torch.float64, device=su.dev)
        for c0 in range(0, B, CH):
            st = f(su, y[c0:c0 + CH], **kw)
            errs_dev += E.count_errors(st.ldpc_idx[-1] if st.ldpc_idx else st.amp_idx[-1], tx[c0:c0 + CH]).sum()
            for a in st.amp_exec:
                its_dev += a.sum()
            del st
        if world > 1:
            dist.all_reduce(errs_dev)                           # the path's collective: error counters
    e1.record()
    torch.cuda.synchronize()
    ms = _reduce([e0.elapsed_time(e1) / reps], world, "max")[0]
    its = _reduce([float(its_dev)], world)[0]
    errs = float(errs_dev)
    bytes_it = (2 * L * M + 3 * su.n) * 8
    info = su.total_bits - (su.nl - su.kl)
    gbps = its / world * bytes_it / (ms / 1e3) / 1e9
    rec = {"shape": name, "note": note, "L": L, "M": M, "n": su.n, "P": P, "flow": flow, "ldpc": lp, "EbN0_ref_dB": db,
           "sigma": sigma, "n_gpus": world, "scaling": scaling, "codewords": total, "ms": ms,
           "codewords_per_s": total / (ms / 1e3), "info_mbit_per_s": total * info / (ms / 1e3) / 1e6,
           "amp_iterations_per_codeword": its / total, "us_per_codeword_iteration": 1e3 * ms * world / max(its, 1),
           "algorithmic_GBps_per_gpu": gbps, "frac": gbps / peak, "final_ber": errs / (total * su.total_bits),
           "amp_mode": amp_mode, "bp_mode": bp_mode}
    if pa:
        rec["power_allocation"] = {"a": pa["a"], "f": pa["f"], "C": pa["C"], "P0_over_PL": float(su.Pl[0] / su.Pl[-1])}
    return rec


def run_exit_point(rank, world, peak, repeats, amp_mode):
    """One point of BASELINE configs[3]: amp_exit_curve's inner loop at L=256, M=32, P=4, 350 bins, `repeats`
    codewords split over the ranks, histograms all-reduced (amp_exit.py:560-595)."""
    from sparc_ldpc_b200 import amp_exit as AX, engine as E, sparc_ldpc as S
    E.AMP_MODE = amp_mode
    group = None
    if world > 1:
        import torch.distributed as dist
        group = dist.group.WORLD
    sp = S.SPARCParams(L=256, M=32, sigma=None, p=4.0, r=1, t=64)
    np.random.seed(11)
    xp = 10
    kw = dict(low_snr_dB=10.0, high_snr_dB=13.0, repeats=repeats, x_axis_points=xp, threshold=0.85, bin_number=350, group=group)
    AX.amp_exit_curve(sp, **dict(kw, repeats=min(repeats, 2 * world)))       # warm-up (tables, allocator)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    _, I_e, _ = AX.amp_exit_curve(sp, **kw)
    e1.record()
    torch.cuda.synchronize()
    ms = _reduce([e0.elapsed_time(e1)], world, "max")[0]
    samples = repeats * 4 * xp                                  # 4 SNR curves x x_axis_points a-priori informations
    return {"shape": "C4-exit", "note": "configs[3]: amp_exit_curve(L=256, M=32, P=4, 4 SNRs in 10..13 dB, x_axis_points=%d, "
                                        "threshold 0.85, 350 bins, repeats=%d): samples split over the ranks, per-sample I_e "
                                        "all-reduced; includes the host-side reference-order RNG walk" % (xp, repeats),
            "n_gpus": world, "scaling": "strong", "repeats": repeats, "codewords": samples, "ms": ms,
            "codewords_per_s": samples / (ms / 1e3), "I_e_mean_per_curve": [float(v) for v in np.mean(I_e, axis=1)],
            "amp_mode": amp_mode}


def run_all(rank, world, dev, peak, quick=True, batch=1184, reps=2, amp_mode="fast", bp_mode="fast", only=None):
    names = [n for n in (QUICK if quick else SHAPES) if not only or n in only]
    out = []
    for name in names:
        out.append(run_shape(name, rank, world, peak, batch, reps, amp_mode, bp_mode))
    if not only or "C4-exit" in only:
        try:
            out.append(run_exit_point(rank, world, peak, 200, amp_mode))
        except Exception as ex:
            out.append({"shape": "C4-exit", "failed": repr(ex)})
    keep = ("shape", "note", "n_gpus", "scaling", "codewords", "ms", "codewords_per_s", "amp_iterations_per_codeword",
            "us_per_codeword_iteration", "frac", "final_ber", "power_allocation", "repeats", "I_e_mean_per_curve", "failed")
    return [{k: r[k] for k in keep if k in r} for r in out] if quick else out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=1184)
    ap.add_argument("--reps", type=int, default=2)
    ap.add_argument("--only", default="")
    ap.add_argument("--amp-mode", default="fast")
    ap.add_argument("--bp-mode", default="fast")
    args = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        peak = 6650.0
    only = [s for s in args.only.split(",") if s]
    for rec in run_all(rank, world, torch.device("cuda", local), peak, quick=False, batch=args.batch, reps=args.reps,
                       amp_mode=args.amp_mode, bp_mode=args.bp_mode, only=only):
        if rank == 0:
            print(json.dumps(rec))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
