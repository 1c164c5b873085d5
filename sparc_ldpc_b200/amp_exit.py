"""Reference-facing EXIT-chart API (ldpc/amp_exit.py): J / J_inverse, gen_bits, hard_initialisation, prep_y,
calc_E, hist_E, calc_I_e, polynomial, amp_exit_curve -- same names, arguments, RNG draw order
(randint(L logM) -> randn(L logM) -> randn(n, 1)) and averaging (per-repetition I_e,
amp_exit.py:590-595); the per-sample work (bp2sp, peel, AMP on the active sections, sp2bp, LLR, clip,
histograms) runs on the device, batched over the (repeat, snr, I_a) triples.
"""
import time

import numpy as np
import torch

from . import decoder as D
from . import engine as E
from . import sparc_ldpc as S
from .sparc_ldpc import SPARCParams, bits2indices, pa_parameterised  # noqa: F401  (re-exported like `import *`)

F64, I32 = torch.float64, torch.int32


def J_inverse(I):
    """amp_exit.py:28-36 (I = 1 is clipped to 0.9999)."""
    assert 0 <= I <= 1
    if I == 1:
        I = np.clip(I, a_min=None, a_max=0.9999)
        print("Warning clipping I from 1 to 0.9999")
    if I <= 0.3646:
        return 1.09542 * (I ** 2) + 0.214217 * I + 2.33727 * np.sqrt(I)
    return -0.706692 * np.log(0.386013 * (1 - I)) + 1.75017 * I


def J(sigma):
    """amp_exit.py:38-45."""
    assert sigma >= 0
    if sigma <= 1.6363:
        return -0.0421061 * (sigma ** 3) + 0.209252 * (sigma ** 2) + -0.00640081 * sigma
    if sigma < 10:
        return 1 - np.exp(0.00181491 * (sigma ** 3) - 0.142675 * (sigma ** 2) - 0.0822054 * sigma + 0.0549608)
    return 1


def gen_bits(length, rng=None):
    """+-1 symbols, +1 <-> bit 0 (amp_exit.py:48-50)."""
    return (S._rng(rng).randint(0, 2, length) * -2) + 1


def _setup(sparcparams):
    return D.make_setup(sparcparams, None)


def prep_y(X, L, M, n, sigma_w, P, a=None, f=None, C=None, rng=None):
    """Channel output for the +-1 message X (amp_exit.py:125-160) -> (y, Ab, Az, Pl, ordering)."""
    Pl = P / L * np.ones(L) if a is None else pa_parameterised(L, C, P, a, f)
    idx = np.asarray(bits2indices((np.asarray(X) - 1) * -1 / 2, M), dtype=np.int32).reshape(1, L)
    Ab, Az, ordering = S.sparc_transforms(L, M, n)
    op = Ab._sb_op
    dev = E._dev()
    x = op.onehot_apply(torch.from_numpy(idx).to(dev), torch.from_numpy(Pl).to(dev)).cpu().numpy().reshape(-1, 1)
    w = S._rng(rng).randn(n, 1) * sigma_w
    return (x + w).reshape(-1, 1), Ab, Az, Pl, ordering


def hard_initialisation(beta, L, M, n, ordering, y, Pl, Ab, threshold=0.5, ldpc_sections=None):
    """Threshold peel (amp_exit.py:56-122).  Like the reference it overwrites `beta` (hard sections ->
    one-hot * sqrt(n Pl), others -> 0) and returns (y_new, Ab_new, Az_new, amp_sections, L_amp_sections)."""
    if ldpc_sections is None:
        ldpc_sections = L
    dev = E._dev()
    op = Ab._sb_op
    post = S._to_dev(beta, L * M)
    hard, act, nact = E.threshold_peel(post, L, M, ldpc_sections, threshold)
    Pld = torch.from_numpy(np.ascontiguousarray(Pl, dtype=np.float64)).to(dev)
    y_new = op.onehot_apply(hard, Pld, S._to_dev(y, n), sign=-1.0).cpu().numpy().reshape(-1, 1)
    beta[:] = E.onehot_beta(hard, Pld, n, L, M).cpu().numpy().reshape(beta.shape)
    La = int(nact[0])
    amp_sections = act[0, :La].cpu().numpy().tolist()
    if La > 0:
        Ab_new, Az_new = S.sparc_transforms_shorter(La, M, n, np.asarray(ordering)[amp_sections, :])
    else:
        Ab_new = Az_new = None
    return y_new, Ab_new, Az_new, amp_sections, La


def calc_E_batch(X, I_a, snr_dB, sparcparams, threshold=0.5, rng=None):
    """B EXIT samples at once.  X: [B, L logM] of +-1; I_a, snr_dB: scalars or [B].  Draw order per sample:
    randn(L logM) for the a-priori LLRs, then randn(n, 1) for the channel (amp_exit.py:218, :157)."""
    rng = S._rng(rng)
    su = _setup(sparcparams)
    X = np.atleast_2d(np.asarray(X))
    B = X.shape[0]
    I_a = np.broadcast_to(np.asarray(I_a, dtype=float), (B,))
    snr_dB = np.broadcast_to(np.asarray(snr_dB, dtype=float), (B,))
    A = np.empty((B, su.total_bits))
    idx = np.empty((B, su.L), dtype=np.int32)
    noise = np.empty((B, su.n))
    for b in range(B):
        sigma_w = np.sqrt(sparcparams.p / 10 ** (snr_dB[b] / 20))      # :205-206 (20 log10 convention)
        sigma_a = J_inverse(I_a[b])
        A[b] = (sigma_a ** 2) / 2 * X[b] + rng.randn(su.total_bits) * sigma_a   # :214-221
        idx[b] = bits2indices((X[b] - 1) * -1 / 2, su.M)
        noise[b] = (rng.randn(su.n, 1) * sigma_w).reshape(-1)
    _, y = S._transmit(su, idx, noise)
    Eo, st = D.exit_E(su, y, torch.from_numpy(A).to(su.dev), threshold)
    return Eo, st


def _full(a):
    """An array printed in full, as the reference's np.set_printoptions(threshold=nan) intends (amp_exit.py:261)."""
    import sys
    return np.array2string(np.asarray(a), threshold=sys.maxsize)


def _export_row(csv_filename, I_a, snr_dB, X, E_host):
    """One (header, row) pair appended per sample, exactly the reference's layout (amp_exit.py:262-268)."""
    import csv
    with open(csv_filename, "a") as fh:
        w = csv.DictWriter(fh, fieldnames=["I_a", "snr_dB", "X", "E"])
        w.writeheader()
        w.writerow({"I_a": I_a, "snr_dB": snr_dB, "X": _full(X), "E": _full(E_host)})


def calc_E(X, I_a, snr_dB, sparcparams, csv_filename=None, threshold=0.5):
    """Extrinsic LLRs of one EXIT sample (amp_exit.py:185-270)."""
    Eo, _ = calc_E_batch(np.asarray(X).reshape(1, -1), I_a, snr_dB, sparcparams, threshold)
    E_host = Eo.cpu().numpy().reshape(-1)
    if csv_filename is not None:
        _export_row(csv_filename, I_a, snr_dB, X, E_host)
    return E_host


class imported_E:
    """E of one sample read back from a CSV export (amp_exit.py:14-25)."""

    def __init__(self, X, E, I_a, SNR_dB):
        self.X, self.E, self.I_a, self.SNR_dB = X, E, I_a, SNR_dB

    def __eq__(self, other):
        return (self.X == other.X).all() and (self.E == other.E).all() and self.I_a == other.I_a and self.SNR_dB == other.SNR_dB


def import_E_fromfile(fileName, datapoints, repeats, Llogm):
    """Reads the (header, row) pairs calc_E(csv_filename=...) appended (amp_exit.py:353-398) into a dictionary keyed
    'round(I_a, 1) round(snr_dB) l' with l the first repetition index still free -- the reference's keys, including
    their rounding (ten I_a values per curve are assumed, :374)."""
    import csv
    import re
    rx = re.compile(r"[-+]? (?: (?: \d* \. \d+ ) | (?: \d+ \.? ) )(?: [Ee] [+-]? \d+ ) ?", re.VERBOSE)
    out = {}
    with open(fileName) as fh:
        for _i in range(repeats):
            for _k in range(10):
                for _j in range(datapoints):
                    row = next(csv.DictReader(fh), None)          # a fresh reader per sample: header, then one row
                    if row is None:
                        return out
                    Ev = np.asarray(rx.findall(row["E"]), dtype=np.float64)
                    Xv = np.asarray(rx.findall(row["X"]), dtype=np.float64).astype(np.int64)
                    I_a, snr = float(row["I_a"]), float(row["snr_dB"])
                    item = imported_E(X=Xv, E=Ev, I_a=I_a, SNR_dB=snr)
                    for l in range(repeats):
                        key = str(np.round(I_a, 1)) + " " + str(int(np.round(snr))) + " " + str(int(l))
                        if key in out:
                            if out[key] == item:
                                print("repeat dict")
                                break
                            continue
                        out[key] = item
                        break
    return out


def _density(counts, edges):
    """np.histogram(..., density=True) from integer counts: n / diff(edges) / n.sum()."""
    counts = counts.astype(float)
    return counts / np.diff(edges) / counts.sum()


def hist_E_batch(X, Eo, bin_number=500, max_bin=40, min_bin=-40):
    """Device histograms for a batch -> (PE_pos [B, nb], PE_neg [B, nb], bin_width)."""
    edges = np.linspace(min_bin, max_bin, bin_number)
    dev = Eo.device
    Xd = X if torch.is_tensor(X) else torch.from_numpy(np.ascontiguousarray(X, dtype=np.int32)).to(dev)
    counts = E.exit_hist(Eo, Xd, torch.from_numpy(edges).to(dev)).cpu().numpy()
    pos = np.stack([_density(c[0], edges) for c in counts])
    neg = np.stack([_density(c[1], edges) for c in counts])
    return pos, neg, (max_bin - min_bin) / (bin_number - 1)


def hist_E(X, E_llr, bin_number=500, max_bin=40, min_bin=-40, plot=False, snr_dB="Not given"):
    """amp_exit.py:272-326 -> (PE_pos, PE_neg, mean_pos, mean_neg, var_pos, var_neg, bin_width)."""
    assert len(E_llr) == len(X)
    dev = E._dev()
    Ed = torch.from_numpy(np.ascontiguousarray(E_llr, dtype=np.float64).reshape(1, -1)).to(dev)
    pos, neg, bw = hist_E_batch(np.asarray(X).reshape(1, -1), Ed, bin_number, max_bin, min_bin)
    PE_pos, PE_neg = pos[0], neg[0]
    edges = np.linspace(min_bin, max_bin, bin_number)
    mids = 0.5 * (edges[1:] + edges[:-1])
    mean_pos = np.average(mids, weights=PE_pos)
    mean_neg = np.average(mids, weights=PE_neg)
    var_pos = np.average((mids - mean_pos) ** 2, weights=PE_pos)
    var_neg = np.average((mids - mean_neg) ** 2, weights=PE_neg)
    return PE_pos, PE_neg, mean_pos, mean_neg, var_pos, var_neg, bw


def calc_I_e(PE_pos, PE_neg, bin_width):
    """Extrinsic mutual information from the two densities (amp_exit.py:328-351)."""
    keep = ~((PE_pos == 0) & (PE_neg == 0))
    PE_pos, PE_neg = PE_pos[keep], PE_neg[keep]
    with np.errstate(divide="ignore", invalid="ignore"):
        i_neg = PE_neg * np.log2(2 * PE_neg / (PE_neg + PE_pos))
        i_pos = PE_pos * np.log2(2 * PE_pos / (PE_neg + PE_pos))
    i_neg[np.isnan(i_neg)] = 0
    i_pos[np.isnan(i_pos)] = 0
    return 1 / 2 * (bin_width * np.sum(i_neg) + bin_width * np.sum(i_pos))


def polynomial(I_a, I_e):
    """Cubic least-squares fit, constant first (amp_exit.py:400-413)."""
    a = np.stack([np.asarray(I_a) ** i for i in range(4)], axis=1)
    return np.linalg.lstsq(a, I_e, rcond=-1)[0]


def amp_exit_curve(sparcparams, low_snr_dB, high_snr_dB, repeats, x_axis_points, threshold, poly_curve=0,
                   bin_number=500, import_data=False, export_csv_filename=None, import_csv_filename=None,
                   chunk=None, rng=None, group=None):
    """AMP EXIT curves at 4 SNRs (amp_exit.py:520-631).  Samples are generated in the reference's nested order
    (repeat, snr, I_a), decoded in device batches, and I_e is computed per sample and averaged over repeats.
    With `group` (a torch.distributed process group, or True for the world group) sample g is decoded by rank
    g mod world and the per-sample I_e values are exchanged (a sum over disjoint supports), then averaged in the
    reference's order: the curve is bit-identical to the one-rank run on every rank.  Returns (I_a_range, I_e [4, x_axis_points], poly_coeff)."""
    t0 = time.time()
    rng = S._rng(rng)
    L, M = sparcparams.L, sparcparams.M
    nbits = int(L * np.log2(M))
    curves = 4
    I_a_range = np.linspace(0, 0.99, x_axis_points)
    snr_dB = np.linspace(low_snr_dB, high_snr_dB, curves)
    if import_data:   # E was exported by an earlier run: only histograms and I_e are computed (amp_exit.py:550-555,:575-583)
        if import_csv_filename is None:
            print("Please enter a valid csv filename in import_csv_filename")
        d = import_E_fromfile(import_csv_filename, curves, repeats, nbits)
        print(len(d))
        acc = np.zeros((curves, x_axis_points))
        for k in range(repeats):
            items = [d[str(np.round(I_a, 1)) + " " + str(int(np.round(s_dB))) + " " + str(k)]
                     for s_dB in snr_dB for I_a in I_a_range]
            for a in items:
                assert len(a.X) == nbits and len(a.E) == nbits
            Xm = np.stack([a.X for a in items])
            Ed = torch.from_numpy(np.stack([a.E for a in items])).to(E._dev())
            pos, neg, bw = hist_E_batch(Xm, Ed, bin_number, 60, -60)
            for q in range(len(items)):
                acc[q // x_axis_points, q % x_axis_points] += calc_I_e(pos[q], neg[q], bw)
        I_e = acc / repeats
        poly_coeff = polynomial(I_a_range, I_e[poly_curve, :])
        print("The coefficients for the polynomial are: ", poly_coeff)
        print("Wall clock time elapsed: ", time.time() - t0)
        return I_a_range, I_e, poly_coeff
    triples = [(k, j, i) for k in range(repeats) for j in range(curves) for i in range(x_axis_points)]
    chunk = chunk or min(len(triples), 512)
    from . import dist as SD
    rank, world = SD.group_info(group)
    pg = None if group is True else group
    ie = np.zeros((repeats, curves, x_axis_points))      # I_e of every sample; each entry is written by ONE rank
    for c0 in range(0, len(triples), chunk):
        part = triples[c0:c0 + chunk]
        # every rank consumes the whole host stream so that sample g is the same codeword everywhere
        Xs, Ias, snrs, states = [], [], [], []
        for (k, j, i) in part:
            Xs.append(gen_bits(nbits, rng))
            states.append(rng.get_state())
            # placeholder draws keep the stream aligned: randn(L logM) then randn(n, 1) per sample
            rng.randn(nbits); rng.randn(int(L * np.log2(M) / sparcparams.r), 1)
            Ias.append(I_a_range[i]); snrs.append(snr_dB[j])
        end_state = rng.get_state()
        mine = [q for q in range(len(part)) if (c0 + q) % world == rank]
        if mine:
            class _Replay:
                """Replays the per-sample draws from the recorded states (sample order = reference order)."""
                def __init__(self, sts):
                    self.sts, self.q, self.r = sts, -1, np.random.RandomState(0)
                def next_sample(self):
                    self.q += 1
                    self.r.set_state(self.sts[self.q])
                def randn(self, *a):
                    return self.r.randn(*a)
            rep = _Replay([states[q] for q in mine])
            Xm = np.stack([Xs[q] for q in mine])
            Eo = _calc_E_replay(Xm, [Ias[q] for q in mine], [snrs[q] for q in mine], sparcparams, threshold, rep)
            if export_csv_filename is not None:                                 # :262-268, rows of this rank
                Eh = Eo.cpu().numpy()
                fn = export_csv_filename if world == 1 else "%s.rank%d" % (export_csv_filename, rank)  # one file per rank
                for m_, q in enumerate(mine):
                    _export_row(fn, Ias[q], snrs[q], Xm[m_], Eh[m_])
            pos, neg, bw = hist_E_batch(Xm, Eo, bin_number, 60, -60)           # :587
            for m_, q in enumerate(mine):
                k, j, i = part[q]
                ie[k, j, i] = calc_I_e(pos[m_], neg[m_], bw)                     # :590
        rng.set_state(end_state)
    if world > 1:   # disjoint supports: the sum is an exact all-gather (x + 0.0 == x), NCCL over NVLink on GPUs
        import torch.distributed as dist
        t = torch.from_numpy(ie).to(E._dev() if dist.get_backend(pg) == "nccl" else "cpu")
        dist.all_reduce(t, group=pg)
        ie = t.cpu().numpy()
    acc = np.zeros((curves, x_axis_points))
    for k in range(repeats):                                                     # the reference's order of additions (:592)
        acc = acc + ie[k]
    I_e = acc / repeats                                                          # :595
    poly_coeff = polynomial(I_a_range, I_e[poly_curve, :])                       # :597
    print("The coefficients for the polynomial are: ", poly_coeff)
    print("Wall clock time elapsed: ", time.time() - t0)
    return I_a_range, I_e, poly_coeff


def _calc_E_replay(X, I_a, snr_dB, sparcparams, threshold, rep):
    """calc_E_batch with each sample's draws replayed from its own recorded RNG state."""
    su = _setup(sparcparams)
    B = X.shape[0]
    A = np.empty((B, su.total_bits))
    idx = np.empty((B, su.L), dtype=np.int32)
    noise = np.empty((B, su.n))
    for b in range(B):
        rep.next_sample()
        sigma_w = np.sqrt(sparcparams.p / 10 ** (snr_dB[b] / 20))
        sigma_a = J_inverse(I_a[b])
        A[b] = (sigma_a ** 2) / 2 * X[b] + rep.randn(su.total_bits) * sigma_a
        idx[b] = bits2indices((X[b] - 1) * -1 / 2, su.M)
        noise[b] = (rep.randn(su.n, 1) * sigma_w).reshape(-1)
    _, y = S._transmit(su, idx, noise)
    Eo, _ = D.exit_E(su, y, torch.from_numpy(A).to(su.dev), threshold)
    return Eo
