#!/usr/bin/env python
"""Headline benchmark: decoded codewords/s (and info Mbit/s) of the SPARC-AMP + outer-LDPC soft-exchange
decoder at L = M = 512, R = 1, P = 4, IEEE 802.16 rate-5/6 outer code (z = 192), 2 AMP<->BP iterations
(BASELINE.json configs[2], the configuration `metric` is quoted on; it fits one GPU).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--impl ours|reference]

A "step" is one pass of the whole soft flow (AMP -> LLR -> BP -> prior -> AMP -> ... , 3 AMP decodes and
2 BP decodes per codeword, sparc_ldpc.py:636-706) over one batch of B synthetic codewords per GPU.  `value`
is measured with the received vectors y already resident in HBM; `e2e` runs the same step from pinned host
buffers through the public API with the H2D copy of y and the D2H read of the decisions / error counts
inside the timed region.  One process per GPU; codewords are independent, so ranks shard them (weak scaling)
and the only collective is the NCCL all-reduce of the error counters.

`--impl reference` times the reference's CPU algorithm (oracle port: numpy + C w-point FHT + the reference's
own c_ldpc.c when oracle/_ref was built) on all host cores, one codeword per worker per step.
"""
import argparse
import contextlib
import io
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# workload (BASELINE.json configs[2]); Eb/N0 in the reference's 20 log10 convention, inside the waterfall
# (ldpc/EbN0_dBVsBER_waterfallsoft_rep200_LM512p4r1rldpc5_6.csv row 7.667 dB)
L, M, P, R_SPARC, T, SOFT_ITER = 512, 512, 4.0, 1, 64, 2
STD, RATE, Z = "802.16", "5/6", 192
EBN0_DB = 7.667
R_TOTAL = 5.0 / 6.0
N = int(L * np.log2(M) / R_SPARC)
INFO_BITS = int(L * np.log2(M)) - (24 * Z - 20 * Z)      # L logM - (N_ldpc - K_ldpc) = 3840
SIGMA = float(np.sqrt(P / ((10 ** (EBN0_DB / 20)) * 2 * R_TOTAL)))   # sparc_ldpc.py:1184,1199-1200
WORKLOAD = "soft AMP<->LDPC exchange x2, L=M=512 R=1 P=4, 802.16 5/6 z=192, Eb/N0(ref dB)=%.3f" % EBN0_DB
METRIC = "decoded codewords/sec at L=M=512 R=1 P=4 +LDPC5/6 (soft exchange, 2 AMP<->BP iterations)"


def peak_hbm():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------ CPU reference arm
def _cpu_one(seed):
    import warnings
    warnings.simplefilter("ignore")
    from oracle import oracle as orc
    rng = np.random.RandomState(seed)
    sp = orc.SPARCParams(L=L, M=M, sigma=SIGMA, p=P, r=R_SPARC, t=T)
    t0 = time.perf_counter()
    res = orc.soft_amp_ldpc_sim(sp, orc.LDPCParams(STD, RATE, Z), SOFT_ITER, rng=rng)
    return time.perf_counter() - t0, res[0], res[1]


def _ensure_oracle():
    if not os.path.isfile(os.path.join(ROOT, "oracle", "_build", "liboracle.so")):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "_build/liboracle.so"],
                              stdout=subprocess.DEVNULL)


def cpu_run(steps, warmup, cores=None):
    """steps x (one soft-flow codeword per worker, all workers in parallel).  Returns cw/s and details."""
    import multiprocessing as mp
    _ensure_oracle()
    cores = cores or os.cpu_count() or 1
    ctx = mp.get_context("fork")
    with ctx.Pool(cores) as pool:
        for w in range(warmup):
            pool.map(_cpu_one, [10_000 + w * cores + i for i in range(cores)])
        t0 = time.perf_counter()
        per = []
        for s in range(steps):
            per += pool.map(_cpu_one, [20_000 + s * cores + i for i in range(cores)])
        wall = time.perf_counter() - t0
    single = float(np.mean([p[0] for p in per]))
    return dict(value=cores * steps / wall, wall=wall, cores=cores, per_codeword_s=single,
                ms_per_step=1e3 * wall / max(steps, 1))


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    res = cpu_run(args.steps, args.warmup)
    kind = "port"
    line = {
        "impl": "reference", "metric": METRIC, "value": res["value"], "unit": "codewords/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": res["ms_per_step"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "info_mbit_per_s": res["value"] * INFO_BITS / 1e6,
        "config": {"workload": WORKLOAD, "codewords_per_step": res["cores"]},
        "cpu_baseline": {"value": res["value"], "unit": "codewords/s", "cores": res["cores"], "kind": kind,
                         "sample": "%d steps x 1 soft-flow codeword per worker on %d workers (%.1f s per codeword per "
                                   "core); oracle port of the reference algorithm (numpy + C w-point FHT, C BP)"
                                   % (args.steps, res["cores"], res["per_codeword_s"])},
        "e2e": {"value": res["value"], "unit": "codewords/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------ clocks sampler
class Clocks(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.rows = index, threading.Event(), []

    def run(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        while not self.stop_flag.is_set() and not os.environ.get("BENCH_NO_CLOCKS"):
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = [float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[2 + i].lower().startswith("active") for r in self.rows if len(r) > 2 + i)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.rows)}


# ------------------------------------------------------------------------------------------ GPU arm
def _union_ms(ev0, pairs):
    """Time during which at least one of the [start, end] event pairs was open (launches of different streams
    overlap), in ms relative to ev0."""
    iv = sorted((ev0.elapsed_time(a), ev0.elapsed_time(b)) for a, b in pairs)
    tot, cs, ce = 0.0, None, None
    for s_, e_ in iv:
        if ce is None or s_ > ce:
            if ce is not None:
                tot += ce - cs
            cs, ce = s_, e_
        else:
            ce = max(ce, e_)
    if ce is not None:
        tot += ce - cs
    return tot


AMP_DESC = {"fast": "fast (fp64 state / transforms / softmax; z and FHT(beta) gathered from 27-bit fixed-point copies; "
                    "stop at |d tau| <= 2^-27 tau)",
            "f64": "f64 (fp64 throughout, exact-equality stop tau == last_tau; gathers summed in a bank-scheduled order: "
                   "differs from strict by fp64 summation-order noise ~1e-15 per iteration; the facades' default)",
            "strict": "strict (fp64 throughout, reference add order, exact-equality stop tau == last_tau)"}
BP_DESC = {"fast": "fast (fp64 messages; the two log(1+exp(-|x|)) terms of every Lxor in single precision)",
           "strict": "strict (fp64 exp/log as c_ldpc.c:246-247)"}


def gpu_arm(args):
    import ctypes as ct

    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)

    from sparc_ldpc_b200 import _lib, decoder as D, engine as E, sparc_ldpc as S

    B = args.batch
    sp = S.SPARCParams(L=L, M=M, sigma=SIGMA, p=P, r=R_SPARC, t=T)
    su = D.make_setup(sp, S.LDPCParams(STD, RATE, Z))
    assert su.n == N and su.total_bits - (su.nl - su.kl) == INFO_BITS
    lxor_per_bp_iteration = int(sum(3 * (int(d) - 2) for d in su.graph.cdeg))     # c_ldpc.c:294-314 per check node

    # synthetic codewords: valid LDPC codewords + AWGN from a per-rank seeded host stream (reference draw order)
    rng = np.random.RandomState(1000 + rank)
    idx, noise = S._draw(su, B, SIGMA, rng)
    tx = torch.from_numpy(idx).to(dev)
    y_dev = su.op.onehot_apply(tx, su.Pl_dev) + torch.from_numpy(noise).to(dev)
    y_host = y_dev.cpu().pin_memory()
    idx_host = torch.empty((B, L), dtype=torch.int32).pin_memory()
    errs_host = torch.empty((2 * SOFT_ITER + 1, B), dtype=torch.int32).pin_memory()
    torch.cuda.synchronize()

    # every AMP / BP launch is bracketed by CUDA events on its launching stream
    amp_ev, amp_exec, bp_ev, bp_its = [], [], [], []
    orig_amp, orig_bp = su.op.amp, su.graph.bp

    def timed_amp(*a, **k):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r = orig_amp(*a, **k)
        e1.record()
        amp_ev.append((e0, e1))
        amp_exec.append(r.n_exec)
        return r

    def timed_bp(*a, **k):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r = orig_bp(*a, **k)
        e1.record()
        bp_ev.append((e0, e1))
        bp_its.append(r[1])
        return r

    su.op.amp, su.graph.bp = timed_amp, timed_bp

    def clear():
        del amp_ev[:], amp_exec[:], bp_ev[:], bp_its[:]

    # The batch is cut into `--streams` slices, each decoded on its own CUDA stream: a slice's dependent chain
    # (AMP -> BP -> AMP ...) is sequential, but the tail of one slice's launch (codewords that run all 64 AMP /
    # 200 BP iterations) overlaps with the next slice's work instead of idling the other SMs.
    all_streams = [torch.cuda.Stream(device=dev) for _ in range(max(1, min(args.streams, B)))]

    def step_device(y, tx_, nstreams):
        b = y.shape[0]
        S_ = max(1, min(nstreams, b))
        bounds = [(i * b) // S_ for i in range(S_ + 1)]
        main = torch.cuda.current_stream()
        outs = []
        for i in range(S_):
            st_ = all_streams[i]
            st_.wait_stream(main)
            with torch.cuda.stream(st_):
                ys = y[bounds[i]:bounds[i + 1]]
                st = D.soft(su, ys, SOFT_ITER)
                stages = [st.amp_idx[0]]
                for j in range(SOFT_ITER):
                    stages += [st.ldpc_idx[j], st.amp_idx[j + 1]]
                txs = tx_[bounds[i]:bounds[i + 1]]
                errs = torch.stack([E.count_errors(s, txs) for s in stages])      # [5, b] bit errors per stage
                outs.append((st, errs))
        for st_ in all_streams[:S_]:
            main.wait_stream(st_)
        errs = torch.cat([o[1] for o in outs], dim=1)
        totals = errs.sum(dim=1, dtype=torch.int64)
        if world > 1:
            dist.all_reduce(totals)                                           # the path's only collective
        return [o[0] for o in outs], errs, totals

    def step_e2e():
        y = y_host.to(dev, non_blocking=True)
        sts, errs, totals = step_device(y, tx, args.streams)
        idx_host.copy_(torch.cat([s.ldpc_idx[-1] for s in sts]), non_blocking=True)   # final decisions
        errs_host.copy_(errs, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return totals

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    peak, peak_src = peak_hbm()
    bytes_per_iter = (2 * L * M + 3 * N) * 8                    # read beta + write beta + read y, read z, write z (fp64)
    traffic_ratio, traffic_src = None, None
    tpath = os.path.join(ROOT, "profiles", "amp_traffic.json")
    if os.path.isfile(tpath):
        try:
            tj = json.load(open(tpath))
            traffic_ratio, traffic_src = tj, "profiles/amp_traffic.json"
        except Exception:
            traffic_ratio = None

    def run_mode(amp_mode, bp_mode, steps, warm, warm_batch):
        """Times `steps` steps with y resident in HBM, one serialised single-stream step (per-launch durations and
        the kernels' share of a step), and `steps` end-to-end steps from pinned host buffers."""
        E.AMP_MODE, E.BP_MODE = amp_mode, bp_mode
        for _ in range(warm):
            step_device(y_dev[:warm_batch], tx[:warm_batch], args.streams)
        barrier()
        clear()
        clocks = Clocks(local)
        clocks.start()
        launches0 = _lib.launch_count()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        ev0.record()
        last = None
        for _ in range(steps):
            last = step_device(y_dev, tx, args.streams)
        ev1.record()
        barrier()
        launches = _lib.launch_count() - launches0
        ms = ev0.elapsed_time(ev1)
        amp_busy_ms = _union_ms(ev0, amp_ev)
        n_amp_launch = len(amp_ev)
        exec_iters = float(sum(int(m.sum()) for m in amp_exec))
        bp_iters = float(sum(int(t.sum()) for t in bp_its))
        clear()

        # one extra, untimed step on ONE stream: launches serialise, so every event pair is that launch's own
        # duration (what an ncu launch list of this command shows) and the kernels' share of a step is defined
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        s0.record()
        step_device(y_dev, tx, 1)
        s1.record()
        barrier()
        ser_ms = s0.elapsed_time(s1)
        ser_amp = [a.elapsed_time(b) for a, b in amp_ev]
        ser_bp = [a.elapsed_time(b) for a, b in bp_ev]
        ser_amp_iters = float(sum(int(m.sum()) for m in amp_exec))
        ser_bp_iters = float(sum(int(t.sum()) for t in bp_its))
        clear()

        for _ in range(min(2, warm)):
            step_e2e()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            step_e2e()
        e1.record()
        barrier()
        ms_e2e = e0.elapsed_time(e1)
        clocks.stop_flag.set()
        clocks.join(2)
        clear()

        t = torch.tensor([ms, ms_e2e, amp_busy_ms, ser_ms, sum(ser_amp), sum(ser_bp)], dtype=torch.float64, device=dev)
        sums = torch.tensor([exec_iters, bp_iters, ser_amp_iters, ser_bp_iters], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dist.all_reduce(sums)
        ms, ms_e2e, amp_busy_ms, ser_ms, ser_amp_ms, ser_bp_ms = (float(v) for v in t)
        exec_iters, bp_iters, ser_amp_iters, ser_bp_iters = (float(v) / world for v in sums)      # per rank
        total_cw = B * world * steps
        alg_bytes = exec_iters * bytes_per_iter
        achieved = alg_bytes / (amp_busy_ms / 1e3) / 1e9
        n_ser = max(len(ser_amp), 1)
        rule = _lib.SB_BP_SUMPROD2_FAST if bp_mode == "fast" else _lib.SB_BP_SUMPROD2
        pk = ct.c_double(0.0)
        bp_peak = None
        if _lib.lib().sb_bp_lxor_peak(rule, ct.byref(pk)) == 0:
            bp_peak = pk.value
        bp_ach = ser_bp_iters * lxor_per_bp_iteration / (ser_bp_ms / 1e3) if ser_bp_ms > 0 else None
        kname = {"fast": "sb::p2::amp2_kernel<false> (two codewords per CTA, q27 gathers)",
                 "f64": "sb::p2::amp2_kernel<true> (one codeword per CTA, fp64 gathers)",
                 "strict": "sb::amp_kernel<9,1,0,0> (reference add order)"}[amp_mode]
        return {
            "value": total_cw / (ms / 1e3), "e2e": total_cw / (ms_e2e / 1e3), "ms_per_step": ms / steps, "steps": steps,
            "launches": int(launches), "errs": last[2].cpu().numpy(), "clocks": clocks.summary(),
            "mean_amp_iterations_per_decode": exec_iters * world / (3.0 * total_cw),
            "mean_bp_iterations_per_decode": bp_iters * world / (2.0 * total_cw),
            "roofline": {
                "bound": "hbm", "kernel": kname,
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": None, "peak_source": peak_src,
                "algorithmic_bytes_per_codeword_iteration": bytes_per_iter,
                "us_per_codeword_iteration": 1e3 * amp_busy_ms / max(exec_iters, 1),
                "launches_timed": n_amp_launch,
                "kernel_busy_ms": amp_busy_ms, "kernel_busy_fraction_of_timed_region": amp_busy_ms / ms,
                "busy_ms_per_launch": amp_busy_ms / max(n_amp_launch, 1),
                "algorithmic_bytes_per_launch": alg_bytes / max(n_amp_launch, 1),
                "serial_step": {"what": "one untimed step on a single stream: launches do not overlap, each event pair is "
                                        "one launch's own duration", "ms": ser_ms, "amp_launches": len(ser_amp),
                                "avg_launch_ms": ser_amp_ms / n_ser,
                                "algorithmic_bytes_per_launch": ser_amp_iters * bytes_per_iter / n_ser,
                                "achieved": ser_amp_iters * bytes_per_iter / (ser_amp_ms / 1e3) / 1e9,
                                "frac": ser_amp_iters * bytes_per_iter / (ser_amp_ms / 1e3) / 1e9 / peak,
                                "kernel_share_of_step": ser_amp_ms / ser_ms, "bp_share_of_step": ser_bp_ms / ser_ms},
            },
            "roofline_bp": {
                "bound": "fp64/sfu instruction rate", "kernel": "sb::bp_kernel_reg<%s>" % ("SUMPROD2_FAST" if bp_mode == "fast" else "SUMPROD2"),
                "achieved": bp_ach, "peak": bp_peak, "unit": "Lxor/s", "frac": (bp_ach / bp_peak) if (bp_ach and bp_peak) else None,
                "lxor_per_codeword_iteration": lxor_per_bp_iteration, "launches_timed": len(ser_bp),
                "avg_launch_ms": ser_bp_ms / max(len(ser_bp), 1),
                "peak_source": "sb_bp_lxor_peak: register-only loop of the same check-node function on all SMs, measured in this run",
                "note": "launch duration includes the tail of codewords that run all 200 iterations while the other "
                        "CTAs have left (hidden by the stream slices in the timed region)"},
        }

    fast = run_mode(args.amp_mode, args.bp_mode, args.steps, max(args.warmup, 3), B)
    strict = f64 = None
    if not args.no_strict and (args.amp_mode, args.bp_mode) != ("strict", "strict"):
        f64 = run_mode("f64", "strict", max(1, args.strict_steps), 1, max(148, B // 8))
        strict = run_mode("strict", "strict", max(1, args.strict_steps), 1, max(148, B // 8))
    su.op.amp, su.graph.bp = orig_amp, orig_bp

    if traffic_ratio is not None:
        for rec, key in ((fast, args.amp_mode), (f64, "f64"), (strict, "strict")):
            if rec is None or key not in traffic_ratio:
                continue
            tr = traffic_ratio[key]
            rl = rec["roofline"]
            rl["traffic"] = float(tr["ratio"]) * rl["serial_step"]["algorithmic_bytes_per_launch"]
            rl["traffic_source"] = ("dram__bytes_read.sum + dram__bytes_write.sum of one ncu --set full capture of this kernel "
                                    "(%s; %s) = %.3f x that launch's algorithmic bytes, scaled to serial_step's bytes per launch"
                                    % (tr.get("kernel", "?"), tr.get("capture", "?"), float(tr["ratio"])))

    shapes = None
    if rank == 0 or world > 1:
        if not args.no_shapes:
            E.AMP_MODE, E.BP_MODE = args.amp_mode, args.bp_mode
            try:
                sys.path.insert(0, os.path.join(ROOT, "tools"))
                import bench_shapes
                shapes = bench_shapes.run_all(rank, world, dev, peak, quick=True)
            except Exception as ex:  # the shapes block is additional information, never the headline
                shapes = {"failed": repr(ex)}

    if rank == 0:
        nbits = B * world * su.total_bits
        amp_mode, bp_mode = args.amp_mode, args.bp_mode
        line = {
            "metric": METRIC, "value": fast["value"], "unit": "codewords/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": fast["ms_per_step"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None,
            "dtype": ("f64 state/transforms/softmax, q27 fixed-point gathers, f32 BP correction terms"
                      if (amp_mode, bp_mode) == ("fast", "fast") else "f64"),
            "data": "synthetic",
            "info_mbit_per_s": fast["value"] * INFO_BITS / 1e6,
            "config": {"workload": WORKLOAD, "codewords_per_step_per_gpu": B, "sigma": SIGMA, "amp_T": T, "streams": args.streams,
                       "amp_mode": AMP_DESC[amp_mode], "bp_mode": BP_DESC[bp_mode],
                       "l2": "working set %.0f MB of beta per GPU per step exceeds the 126 MB L2" % (B * L * M * 8 / 1e6),
                       "ber_per_stage[amp1,ldpc1,amp2,ldpc2,amp3]": (fast["errs"] / nbits).tolist(),
                       "mean_amp_iterations_per_decode": fast["mean_amp_iterations_per_decode"],
                       "mean_bp_iterations_per_decode": fast["mean_bp_iterations_per_decode"]},
            "e2e": {"value": fast["e2e"], "unit": "codewords/s", "h2d_bytes_per_step": int(B * N * 8),
                    "d2h_bytes_per_step": int(idx_host.numel() * 4 + errs_host.numel() * 4),
                    "info_mbit_per_s": fast["e2e"] * INFO_BITS / 1e6},
            "gpu_launches": fast["launches"],
            "roofline": fast["roofline"], "roofline_bp": fast["roofline_bp"],
            "clocks": fast["clocks"],
        }
        line["roofline"]["co_limiter"] = (
            "not HBM: the L1 / shared-memory data pipe (2 L n random shared-memory reads per codeword-iteration + the "
            "L2-resident table words).  ncu of this build: l1tex__data_pipe_lsu_wavefronts 81 % of peak (F64 kernel 91 %, "
            "STRICT 85 % with half of its wavefronts bank conflicts), dram / algorithmic bytes 0.92; as a roofline of that "
            "pipe: 532 k wavefronts per codeword-iteration at 1 per cycle and SM = 1.84 us, measured 2.31 in an isolated "
            "launch (profiles/r02_amp_kernel_ncu_full.csv, r02_amp2_hotspots.txt, DESIGN.md section 6b)")
        def ref_record(rec, amp_key):
            return {
                "what": "the same step with amp %s + bp strict, %d timed step(s) after 1 warm-up step on %d codewords"
                        % (amp_key, rec["steps"], max(148, B // 8)),
                "amp_mode": AMP_DESC[amp_key], "bp_mode": BP_DESC["strict"], "dtype": "f64",
                "value": rec["value"], "e2e": rec["e2e"], "unit": "codewords/s", "ms_per_step": rec["ms_per_step"],
                "steps": rec["steps"], "gpu_launches": rec["launches"],
                "mean_amp_iterations_per_decode": rec["mean_amp_iterations_per_decode"],
                "mean_bp_iterations_per_decode": rec["mean_bp_iterations_per_decode"],
                "ber_per_stage[amp1,ldpc1,amp2,ldpc2,amp3]": (rec["errs"] / nbits).tolist(),
                "roofline": rec["roofline"], "roofline_bp": rec["roofline_bp"], "clocks": rec["clocks"]}

        if strict is not None:
            # strict = the reference's arithmetic AND add order; f64 = the same arithmetic type and stop rule with
            # scheduled gathers, the mode the reference-facing facades run by default (engine.AMP_MODE)
            line["strict"] = ref_record(strict, "strict")
            line["f64"] = ref_record(f64, "f64")
            line["iterations_ratio_vs_strict"] = (fast["mean_amp_iterations_per_decode"] /
                                                  strict["mean_amp_iterations_per_decode"])
            line["speedup_vs_strict"] = {
                "value": fast["value"] / strict["value"],
                "from_fewer_amp_iterations": strict["mean_amp_iterations_per_decode"] / fast["mean_amp_iterations_per_decode"],
                "from_faster_amp_iteration": strict["roofline"]["us_per_codeword_iteration"] / fast["roofline"]["us_per_codeword_iteration"],
                "f64_over_strict": f64["value"] / strict["value"], "fast_over_f64": fast["value"] / f64["value"]}
        if shapes is not None:
            line["shapes"] = shapes
        if world == 1 and not args.no_cpu:
            try:
                c = cpu_run(1, 0)
                line["cpu_baseline"] = {"value": c["value"], "unit": "codewords/s", "cores": c["cores"], "kind": "port",
                                        "sample": "1 soft-flow codeword per worker on %d workers (%.1f s per codeword per core); "
                                                  "oracle port of the reference algorithm; the unmodified reference needs 2.7x the "
                                                  "port's time for the same codeword with identical BER tuples (build container, "
                                                  "one core each: profiles/r02_port_vs_reference.json)" % (c["cores"], c["per_codeword_s"])}
            except Exception as ex:  # the baseline is reported, never the product path
                line["cpu_baseline"] = {"value": None, "unit": "codewords/s", "cores": 0, "kind": "port", "sample": "failed: %r" % ex}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=9472,
                    help="codewords per step per GPU (64 x 148 SMs; one CTA per codeword, several waves and streams "
                         "even out the per-codeword early stop; 2368 -> 4736 -> 9472 amortises the end-of-step tail: "
                         "+6 %%, +1.3 %%)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--streams", type=int, default=4,
                    help="slices of the batch decoded on concurrent CUDA streams (round-2 sweep at batch 9472: 2 / 3 / 4 / 8 / 16 "
                         "streams -> 6061 / 6071 / 6090 / 6039 / 5935 codewords/s: longer launches keep both slots of the pair "
                         "kernel's CTAs filled, a few slices still hide the launches' tails)")
    ap.add_argument("--no-strict", action="store_true", help="skip the strict/strict record")
    ap.add_argument("--strict-steps", type=int, default=1, help="timed steps of the strict/strict record")
    ap.add_argument("--no-shapes", action="store_true", help="skip the per-shape block (BASELINE.json configs)")
    ap.add_argument("--amp-mode", default="fast", choices=["strict", "fast"],
                    help="AMP arithmetic: strict = fp64 in the reference's add order; fast = fp64 with 32-bit "
                         "fixed-point gathers (include/sparc_b200.h SB_AMP_FAST); both pass the parity tests")
    ap.add_argument("--bp-mode", default="fast", choices=["strict", "fast"],
                    help="sum-product check nodes: strict = fp64 exp/log; fast = single-precision Lxor correction "
                         "terms (SB_BP_SUMPROD2_FAST); decisions of convergent blocks are identical (tests)")
    args = ap.parse_args()
    # stdout carries exactly ONE JSON line: anything a library prints there meanwhile (NCCL's version banner under
    # torchrun, for one) is sent to stderr, and the real stdout is restored for the result line only
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    out = io.StringIO()
    with contextlib.redirect_stdout(out):
        if args.impl == "reference":
            reference_arm(args)
        else:
            gpu_arm(args)
    sys.stdout.flush()
    os.dup2(real_stdout, 1)
    os.close(real_stdout)
    lines = [ln for ln in out.getvalue().splitlines() if ln.strip()]
    result = lines[-1] if (lines and lines[-1].lstrip().startswith("{")) else None   # rank 0's JSON line; other ranks have none
    for ln in (lines[:-1] if result else lines):
        print(ln, file=sys.stderr)
    if result:
        print(result, flush=True)


if __name__ == "__main__":
    main()
