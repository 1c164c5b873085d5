"""Throughput-mode Monte-Carlo points (BASELINE configs[4]): codewords are generated ON
THE DEVICE -- information bits and channel noise from torch's Philox generator, LDPC encoding and the SPARC
encoder by libsparc_b200 kernels -- so that 10k codewords per Eb/N0 point do not wait for the host's numpy
stream.  This mode is explicitly NOT stream-compatible with the reference (parity runs use the host draws of
sparc_ldpc.py); what it shares with them is every decoder kernel.

Global codeword g is decoded by rank g mod world; the only collective is the all-reduce of the counters.
"""
import numpy as np
import torch

from . import decoder as D
from . import dist as SD
from . import engine as E

F64, I32, U8 = torch.float64, torch.int32, torch.uint8

FLOWS = {"plain": D.plain, "soft": D.soft, "hard": D.hard_init, "originalHard": D.original_hard, "threshold": D.threshold}


def generate(su, B, sigma, gen):
    """B codewords on the device -> (tx_idx [B, L] int32, y [B, n] float64)."""
    dev = su.dev
    if su.code is None:
        bits = torch.randint(0, 2, (B, su.total_bits), dtype=U8, device=dev, generator=gen)
    else:
        info = torch.randint(0, 2, (B, su.kl), dtype=U8, device=dev, generator=gen)
        cw = E.ldpc_encode(su.code, info)                                   # ldpc.py:790-850
        free = torch.randint(0, 2, (B, su.total_bits - su.nl), dtype=U8, device=dev, generator=gen)
        bits = torch.cat([free, cw], dim=1).contiguous()                    # unprotected bits first (sparc_ldpc.py:426)
    tx = E.bits2idx(bits, su.L, su.M)
    noise = torch.randn((B, su.n), dtype=F64, device=dev, generator=gen) * sigma
    y = su.op.onehot_apply(tx, su.Pl_dev) + noise
    return tx, y


def ber_point(sparcparams, ldpcparams, n_codewords, flow="soft", soft_iter=2, threshold=0.6, seed=0, batch=1184,
              amp_mode=None, group=None):
    """One Eb/N0 point of `n_codewords` device-generated codewords.  Returns a dict with the per-stage BER in the
    order the reference reports them (AMP stages, then LDPC stages), block-error counts and iteration totals,
    summed over all ranks of `group`."""
    if flow not in FLOWS:
        raise ValueError("flow must be one of %s" % sorted(FLOWS))
    su = D.make_setup(sparcparams, None if flow == "plain" else ldpcparams)
    rank, world = SD.world()
    mine = SD.shard_indices(n_codewords, rank, world)
    gen = torch.Generator(device=su.dev)
    gen.manual_seed(int(seed) * 1_000_003 + rank)
    kw = {}
    if flow == "soft":
        kw = dict(soft_iter=soft_iter)
    elif flow == "threshold":
        kw = dict(soft_iter=soft_iter, thr=threshold)
    prev_mode = E.AMP_MODE
    if amp_mode is not None:
        E.AMP_MODE = amp_mode
    totals = None
    try:
        for b0 in range(0, len(mine), batch):
            B = min(batch, len(mine) - b0)
            tx, y = generate(su, B, sparcparams.sigma, gen)
            st = FLOWS[flow](su, y, **kw)
            stages = st.amp_idx + st.ldpc_idx
            errs = torch.stack([E.count_errors(s, tx) for s in stages]).to(torch.int64)       # [stages, B]
            row = torch.cat([errs.sum(dim=1), (errs > 0).sum(dim=1),
                             torch.stack([a.sum() for a in st.amp_exec]).to(torch.int64).sum().reshape(1),
                             (torch.stack([t.sum() for t in st.bp_it]).to(torch.int64).sum().reshape(1)
                              if st.bp_it else torch.zeros(1, dtype=torch.int64, device=su.dev)),
                             st.ref_nan_count().to(su.dev).reshape(1), torch.tensor([B], dtype=torch.int64, device=su.dev)])
            totals = row if totals is None else totals + row
            n_amp, n_ldpc = len(st.amp_idx), len(st.ldpc_idx)
    finally:
        E.AMP_MODE = prev_mode
    if totals is None:  # this rank had no codeword: still take part in the reduction
        n_amp, n_ldpc = _stage_counts(flow, soft_iter, su)
        totals = torch.zeros(2 * (n_amp + n_ldpc) + 4, dtype=torch.int64, device=su.dev)
    tot = SD.allreduce_counts(totals, group)
    ns = n_amp + n_ldpc
    n = int(tot[-1])
    bits = n * su.total_bits
    return dict(n_codewords=n, R=su.R, ber_amp=(tot[:n_amp] / bits).tolist(), ber_ldpc=(tot[n_amp:ns] / bits).tolist(),
                block_errors_amp=tot[ns:ns + n_amp].tolist(), block_errors_ldpc=tot[ns + n_amp:2 * ns].tolist(),
                amp_iterations=int(tot[2 * ns]), bp_iterations=int(tot[2 * ns + 1]), amp_ref_nan=int(tot[2 * ns + 2]))


def _stage_counts(flow, soft_iter, su):
    if flow == "plain":
        return 1, 0
    if flow == "soft":
        return soft_iter + 1, soft_iter
    if flow == "hard":
        return 2, 1
    if flow == "originalHard":
        return (2 if su.L - su.ls > 0 else 1), 1
    return soft_iter, soft_iter


def waterfall_device(sparcparams, ldpcparams, EbN0_dB, n_codewords, R=None, flow="soft", **kw):
    """BER-vs-Eb/N0 sweep in throughput mode with the reference's sigma convention
    (ebno = 10^(dB/20), snr = 2 R ebno, sigma = sqrt(P/snr); sparc_ldpc.py:1184,1199-1200)."""
    from .sparc_ldpc import SPARCParams
    sp = sparcparams
    out = []
    dbs = np.asarray(EbN0_dB, dtype=float)
    base_seed = int(kw.pop("seed", 0))          # read ONCE: every point of a sweep derives its stream from it
    for i, db in enumerate(dbs):
        su = D.make_setup(sp, None if flow == "plain" else ldpcparams)
        Rr = su.R if R is None else R
        sigma = float(np.sqrt(sp.p / (10 ** (db / 20) * 2 * Rr)))
        res = ber_point(SPARCParams(sp.L, sp.M, sigma, sp.p, sp.r, sp.t, sp.a, sp.f, sp.C), ldpcparams, n_codewords,
                        flow=flow, seed=base_seed * len(dbs) + i, **kw)
        res["EbN0_dB"], res["sigma"] = float(db), sigma
        out.append(res)
    return out
