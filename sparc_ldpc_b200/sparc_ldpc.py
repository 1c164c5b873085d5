"""Reference-facing host API: the entry points of ldpc/sparc_ldpc.py with the same names, argument meaning,
return values, RNG draw order, stop rules and CSV schemas -- and the per-codeword work done by
libsparc_b200 on the GPU.  Thin host code only: parameter bookkeeping, numpy RNG draws (legacy global
stream, reference order), LDPC encoding, Monte-Carlo stop-rule replay, CSV writing.

Per-codeword functions keep the reference signature and gain a `*_batch` twin that decodes B codewords in
one pass; the single-codeword form is the batch form with B = 1.  There is no CPU fallback: without the
CUDA library these functions raise.
"""
import csv
import math
import os

import numpy as np
import torch

from . import decoder as D
from . import dist as SD
from . import engine as E
from .ldpc import get_code

F64, I32 = torch.float64, torch.int32


# ----------------------------------------------------------------------------------- parameter objects
class SPARCParams:
    """sparc_ldpc.py:227-246."""

    def __init__(self, L, M, sigma, p, r, t, a=None, f=None, C=None):
        self.L, self.M, self.sigma, self.p, self.r, self.t = L, M, sigma, p, r, t
        self.a, self.f, self.C = a, f, C


class LDPCParams:
    """sparc_ldpc.py:250-255."""

    def __init__(self, standard, r_ldpc, z, ptype="A"):
        self.standard, self.r_ldpc, self.z, self.ptype = standard, r_ldpc, z, ptype


def pa_parameterised(L, C, P, a, f):
    """Exponential power allocation, flat from int(f*L) on, scaled to sum P (sparc_ldpc.py:172-186).
    Raises IndexError when f >= 1 exactly like the reference (:184)."""
    pa = 2 ** (-2 * a * C * np.arange(L) / L)
    pa[int(f * L):] = pa[int(f * L)]
    pa /= pa.sum() / P
    return pa


# ----------------------------------------------------------------------------------- operators
def fht_inplace(x):
    """pyfht.fht_inplace (sparc_ldpc.py:14-29): in-place unnormalised Walsh-Hadamard transform of a 1-D
    C-contiguous float64 array whose length is a power of two; same butterfly order, hence bit-identical."""
    if not (isinstance(x, np.ndarray) and x.dtype == np.float64 and x.flags.c_contiguous and x.ndim == 1):
        raise TypeError("fht_inplace needs a 1-D C-contiguous float64 array")
    from . import _lib
    _lib.check(_lib.lib().sb_fht_inplace_host(x.ctypes.data, x.size), "sb_fht_inplace_host")


def _to_dev(a, cols):
    t = torch.from_numpy(np.ascontiguousarray(np.asarray(a, dtype=np.float64).reshape(1, -1))).cuda()
    if t.shape[1] != cols:
        raise AssertionError("operand has %d entries, expected %d" % (t.shape[1], cols))
    return t


def _closures(op, L_use, rows):
    """Ab/Az closures over an Operator restricted to the section list `rows` (None = all)."""
    dev = E._dev()
    if rows is None:
        sections = nsec = None
    else:
        sec = np.zeros((1, op.L), dtype=np.int32)
        sec[0, :L_use] = rows
        sections = torch.from_numpy(sec).to(dev)
        nsec = torch.tensor([L_use], dtype=I32, device=dev)

    def Ab(b):
        b = np.asarray(b, dtype=np.float64).reshape(-1)
        assert b.size == L_use * op.M
        full = np.zeros(op.L * op.M)
        full[:b.size] = b
        return op.Ab(_to_dev(full, op.L * op.M), sections, nsec).cpu().numpy().reshape(-1, 1)

    def Az(z):
        z = np.asarray(z, dtype=np.float64).reshape(-1)
        assert z.size == op.n
        return op.Az(_to_dev(z, op.n), sections, nsec).cpu().numpy().reshape(-1)[:L_use * op.M].reshape(-1, 1)

    for f in (Ab, Az):
        f._sb_op, f._sb_rows, f._sb_L = op, rows, L_use
    return Ab, Az


class _Ordering(np.ndarray):
    """The (L, n) uint32 ordering array, remembering which device operator it belongs to so that
    sparc_transforms_shorter(ordering[rows]) can reuse the tables instead of rebuilding them."""

    def __array_finalize__(self, obj):
        self._sb_op = getattr(obj, "_sb_op", None)


def sparc_transforms(L, M, n, seed=0):
    """(Ab, Az, ordering) of sparc_ldpc.py:140-147; Ab: (LM,)|(LM,1) -> (n,1), Az: (n,)|(n,1) -> (LM,1)."""
    op = E.get_operator(L, M, n, seed)
    Ab, Az = _closures(op, L, None)
    ordering = op.ordering.view(_Ordering)
    ordering._sb_op = op
    return Ab, Az, ordering


def sparc_transforms_shorter(L, M, n, ordering):
    """Same operator restricted to the first L rows of `ordering` (sparc_ldpc.py:154-168).  Callers pass the
    leading block (:522) or a fancy-indexed active set (amp_exit.py:113-116); both are rows of a known table."""
    ordering = np.asarray(ordering)[:L]
    op = E.Operator(L, M, n, ordering=ordering)
    return _closures(op, L, None)


def sparc_transforms_gaussian(L, M, n, seed=0, A=None):
    """Dense i.i.d. Gaussian design matrix A = RandomState(seed).randn(n, L*M)/sqrt(n) (BASELINE config 1;
    the reference has no such matrix, but its amp() takes any closures) -> (Ab, Az, A)."""
    if A is None:
        A = np.random.RandomState(seed).randn(n, L * M) / np.sqrt(n)
    A = np.ascontiguousarray(A, dtype=np.float64)
    op = E.DenseOperator(A, L, M)

    def Ab(b):
        return op.Ab(_to_dev(b, L * M)).cpu().numpy().reshape(-1, 1)

    def Az(z):
        return op.Az(_to_dev(z, n)).cpu().numpy().reshape(-1, 1)

    for f in (Ab, Az):
        f._sb_op, f._sb_rows, f._sb_L = op, None, L
    return Ab, Az, A


def amp(y, sigma_n, Pl, L, M, T, Ab, Az, beta=None):
    """AMP decoder with the reference signature (sparc_ldpc.py:189-222); sigma_n is unused, as there.
    With the closures returned by sparc_transforms[_shorter] / sparc_transforms_gaussian of this module the whole
    loop is one device kernel; with ANY other Ab / Az callables (the reference accepts any, :189) the loop runs
    here, the two products are the caller's closures and the section softmax is the device kernel
    sb_section_softmax_batch."""
    return _amp_host(y, Pl, L, M, T, Ab, Az, beta)[0]


def _amp_foreign(y, Pl, L, M, T, Ab, Az, beta):
    """sparc_ldpc.py:189-222 line for line around foreign closures; returns (beta (LM,1), t)."""
    Pl = np.ascontiguousarray(Pl, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64).reshape(-1, 1)
    P, n = np.sum(Pl), y.size                                   # (:190-191)
    no_init = beta is None or (isinstance(beta, np.ndarray) and beta.dtype == object)
    if no_init:
        b, z = np.zeros((L * M, 1)), y                          # (:193-195)
    else:
        b = np.asarray(beta, dtype=np.float64).reshape(L * M, 1)
        z = y - np.asarray(Ab(b), dtype=np.float64).reshape(-1, 1)   # (:197-198)
    last_tau, t = 0, 0
    Pld = torch.from_numpy(Pl).cuda()
    for t in range(T):
        tau = np.sqrt(np.sum(z ** 2) / n)                       # (:203)
        if tau == last_tau:                                     # (:204)
            return b, t
        last_tau = tau
        s = b + np.asarray(Az(z), dtype=np.float64).reshape(-1, 1)   # (:213)
        sd = torch.from_numpy(np.ascontiguousarray(s.reshape(1, -1))).cuda()
        t2 = torch.tensor([tau ** 2], dtype=F64, device=sd.device)
        bd, _ = E.section_softmax(sd, Pld, t2, L, M, n)         # (:214-219) on the device
        b = bd.cpu().numpy().reshape(-1, 1)
        z = y - np.asarray(Ab(b), dtype=np.float64).reshape(-1, 1) + (z / tau ** 2) * (P - np.sum(b ** 2) / n)   # (:220)
    return b, t


def _amp_host(y, Pl, L, M, T, Ab, Az, beta):
    op = getattr(Ab, "_sb_op", None)
    if op is None or getattr(Az, "_sb_op", None) is not op:
        return _amp_foreign(y, Pl, L, M, T, Ab, Az, beta)
    if op.L != L or op.M != M:
        raise AssertionError("operator was built for L=%d, M=%d" % (op.L, op.M))
    no_init = beta is None or (isinstance(beta, np.ndarray) and beta.dtype == object)  # the [None] sentinel (:189)
    yd = _to_dev(y, op.n)
    Pld = torch.from_numpy(np.ascontiguousarray(Pl, dtype=np.float64)).cuda()
    b0 = None if no_init else _to_dev(beta, L * M)
    res = op.amp(yd, Pld, T, beta0=b0)
    return res.beta.cpu().numpy().reshape(-1, 1), int(res.iters[0])


def amp_gaussian_batch(op, y, Pl, T, beta0=None):
    """Batched AMP over a DenseOperator: y [B, n] numpy -> (beta [B, L*M], iters [B])."""
    yd = torch.from_numpy(np.ascontiguousarray(y, dtype=np.float64)).cuda()
    Pld = torch.from_numpy(np.ascontiguousarray(Pl, dtype=np.float64)).cuda()
    b0 = None if beta0 is None else torch.from_numpy(np.ascontiguousarray(beta0, dtype=np.float64)).cuda()
    res = op.amp(yd, Pld, T, beta0=b0)
    return res.beta.cpu().numpy(), res.iters.cpu().numpy()


# ----------------------------------------------------------------------------------- section <-> bit maps
def sp2bp(beta, L, M):
    """Bit-1 marginals, MSB first (sparc_ldpc.py:257-281); beta holds normalised section posteriors."""
    b = _to_dev(beta, L * M)
    ones = torch.full((L,), 1.0, dtype=F64, device=b.device)
    # scale = sqrt(n * Pl) with n = 1, Pl = 1 -> division by exactly 1.0
    _, p = E.sp2bp_llr(b, M, 1, ones, count=L, want_p=True)
    return p.cpu().numpy().reshape(-1)


def bp2sp(v, L, M):
    """Section posteriors from independent bit posteriors (sparc_ldpc.py:283-314)."""
    a = _to_dev(v, L * int(np.log2(M)))
    ones = torch.full((L,), 1.0, dtype=F64, device=a.device)
    return E.bp2sp_prior(a, L, None, L, M, 1, ones, False, from_prob=True).cpu().numpy().reshape(-1)


def bits2indices(bits, m):
    """MSB-first bits -> section indices (sparc_ldpc.py:317-341); plain host integer work."""
    logm = int(math.log(m, 2))
    b = np.asarray(bits).astype(bool).astype(np.int64)
    assert len(b) % logm == 0
    return (b.reshape(-1, logm) << np.arange(logm - 1, -1, -1)).sum(axis=1).tolist()


def ber_from_LLRs(M, LLR, input_indices, total_bits):
    """sparc_ldpc.py:343-356."""
    logm = int(np.log2(M))
    llr = _to_dev(LLR, len(LLR))
    idx = E.llr2idx(llr, len(LLR) // logm, M)
    tx = torch.tensor(np.asarray(input_indices, dtype=np.int32).reshape(1, -1), device=llr.device)
    return int(E.count_errors(idx, tx)[0]) / total_bits


# ----------------------------------------------------------------------------------- codeword generation
def _rng(rng):
    return np.random if rng is None else rng


def _draw(su, B, sigma, rng, all_zero=False):
    """Messages and channel noise for B codewords in the reference's draw order:
    per codeword randint(kl) -> encode -> randint(total_bits - nl) -> randn(n, 1) * sigma."""
    idx = np.empty((B, su.L), dtype=np.int32)
    noise = np.empty((B, su.n))
    for b in range(B):
        if all_zero:
            bits = np.zeros(su.total_bits, dtype=int)            # sparc_ldpc.py:932
        elif su.code is None:
            bits = rng.randint(0, 2, su.total_bits)              # :424 with nl = 0
        else:
            protected = rng.randint(0, 2, su.kl)                 # :419
            ldpc_bits = su.code.encode(protected)                # :421
            unprotected = rng.randint(0, 2, int(su.total_bits - su.nl))  # :424
            bits = np.concatenate([unprotected, ldpc_bits])      # :426
        idx[b] = bits2indices(bits, su.M)
        noise[b] = (rng.randn(su.n, 1) * sigma).reshape(-1)      # :445
    return idx, noise


def _transmit(su, idx, noise):
    tx = torch.from_numpy(idx).to(su.dev)
    x = su.op.onehot_apply(tx, su.Pl_dev)                        # x = A beta_0   (:436-439)
    y = x + torch.from_numpy(noise).to(su.dev)                   # y = x + z      (:446)
    return tx, y


def _matrix_seed(seed, rng):
    """seed=None is the reference's sparc_transforms(seed=None) of the custom-protograph path (sparc_ldpc.py:930-941):
    a design matrix drawn from OS entropy for EVERY codeword.  Here one matrix serves all codewords of a driver call
    (the BER is conditioned on that draw instead of averaged over the ensemble -- a documented deviation: tables for
    one matrix cost ~1 s to build), and its seed is derived from the position of the host RNG stream WITHOUT
    consuming a draw, so that runs can be replayed and all ranks of a group build the same matrix."""
    if seed is not None:
        return seed
    import zlib
    return int(zlib.crc32(_rng(rng).get_state()[1].tobytes()) & 0x7FFFFFFF)


def _run(flow, sparcparams, ldpcparams, B, rng, seed=0, all_zero=False, **kw):
    su = D.make_setup(sparcparams, ldpcparams, seed=_matrix_seed(seed, rng))
    idx, noise = _draw(su, B, sparcparams.sigma, _rng(rng), all_zero)
    tx, y = _transmit(su, idx, noise)
    st = flow(su, y, **kw)
    ber_amp, ber_ldpc = st.ber(tx, su.total_bits)
    return su, st, ber_amp, ber_ldpc


# ----------------------------------------------------------------------------------- link simulations
def amp_ldpc_sim_batch(sparcparams, ldpcparams=None, B=1, rng=None):
    """B codewords of amp_ldpc_sim -> (ber_amp [B], ber_ldpc [B] | None, ber_ldpc_amp [B] | None, R)."""
    if ldpcparams is None:
        su, st, ba, _ = _run(D.plain, sparcparams, None, B, rng)
        return ba[:, 0], None, None, su.R
    su, st, ba, bl = _run(D.original_hard, sparcparams, ldpcparams, B, rng)
    return ba[:, 0], bl[:, 0], (ba[:, 1] if ba.shape[1] > 1 else None), su.R


def amp_ldpc_sim(sparcparams, ldpcparams=None, a=None, f=None, C=None):
    """sparc_ldpc.py:359-545 ("original hard" exchange; plain SPARC when ldpcparams is None)."""
    ba, bl, bla, R = amp_ldpc_sim_batch(sparcparams, ldpcparams, 1)
    return float(ba[0]), None if bl is None else float(bl[0]), None if bla is None else float(bla[0]), R


def soft_amp_ldpc_sim_batch(sparcparams, ldpcparams, soft_iter, B=1, rng=None):
    su, st, ba, bl = _run(D.soft, sparcparams, ldpcparams, B, rng, soft_iter=soft_iter)
    return ba, bl, su.R


def soft_amp_ldpc_sim(sparcparams, ldpcparams, soft_iter, a=None, f=None, C=None):
    """sparc_ldpc.py:547-712 -> (ber_amp[soft_iter+1], ber_ldpc[soft_iter], R)."""
    ba, bl, R = soft_amp_ldpc_sim_batch(sparcparams, ldpcparams, soft_iter, 1)
    return ba[0].tolist(), bl[0].tolist(), R


def hardinitbeta_amp_ldpc_sim_batch(sparcparams, ldpcparams, B=1, rng=None):
    su, st, ba, bl = _run(D.hard_init, sparcparams, ldpcparams, B, rng)
    return ba, bl, su.R


def hardinitbeta_amp_ldpc_sim(sparcparams, ldpcparams):
    """sparc_ldpc.py:715-860 -> (ber_amp[2], ber_ldpc[1], R)."""
    ba, bl, R = hardinitbeta_amp_ldpc_sim_batch(sparcparams, ldpcparams, 1)
    return ba[0].tolist(), bl[0].tolist(), R


def soft_amp_ldpc_hardinit_batch(sparcparams, ldpcparams, soft_iter, threshold, B=1, rng=None):
    ieee = ldpcparams.standard in ("802.11n", "802.16")
    # custom protographs: all-zero message and a fresh design matrix per call (sparc_ldpc.py:930-935; _matrix_seed)
    su, st, ba, bl = _run(D.threshold, sparcparams, ldpcparams, B, rng, seed=0 if ieee else None,
                          all_zero=not ieee, soft_iter=soft_iter, thr=threshold)
    return ba, bl, su.R


def soft_amp_ldpc_hardinit(sparcparams, ldpcparams, soft_iter, threshold):
    """sparc_ldpc.py:862-1046 -> (ber_amp[soft_iter], ber_ldpc[soft_iter], R)."""
    ba, bl, R = soft_amp_ldpc_hardinit_batch(sparcparams, ldpcparams, soft_iter, threshold, 1)
    return ba[0].tolist(), bl[0].tolist(), R


# ----------------------------------------------------------------------------------- BPSK baseline
def awgn(x, sigma):
    return x + sigma * np.random.randn(len(x))      # sparc_ldpc.py:1049-1052


def ch2llr(ch, sigma):
    return 2.0 / sigma ** 2 * ch                    # :1054-1057


def bpsk(x):
    return 1.0 - 2.0 * x                            # :1059-1061


_RATES = ("1/2", "2/3", "3/4", "5/6", "0.45")


def sim_ldpc(ldpcparams, sigma, MIN_ERRORS=100, MAX_BLOCKS=400000, chunk=256, rng=None, group=None):
    """LDPC + BPSK over AWGN (sparc_ldpc.py:1064-1124), decoded in chunks on the device; the sequential
    stop rule (MIN_ERRORS block errors or MAX_BLOCKS) is replayed on the host and the RNG is rewound to
    where the reference would have stopped drawing.  With `group` (torch.distributed process group, or True for
    the world group) the blocks of every chunk are decoded by rank j mod world and the error counts exchanged
    (dist.decode_sharded): same result and RNG state on every rank as the one-rank run."""
    if ldpcparams.r_ldpc not in _RATES:
        raise NameError("Rate unsupported")
    rng = _rng(rng)
    code = get_code(ldpcparams.standard, ldpcparams.r_ldpc, ldpcparams.z, ldpcparams.ptype)
    ieee = ldpcparams.standard in ("802.11n", "802.16")
    dev = E._dev()
    nbit = nblockerr = nblocks = 0
    ber = 0.0
    while nblockerr < MIN_ERRORS and nblocks < MAX_BLOCKS:
        nb = int(min(chunk, MAX_BLOCKS - nblocks))
        states, xs, ys = [], [], []
        for _ in range(nb):
            states.append(rng.get_state())
            x = code.encode(rng.randint(0, 2, code.K)) if ieee else np.zeros(code.N)
            xs.append(x)
            ys.append(ch2llr(bpsk(x) + sigma * rng.randn(len(x)), sigma))
        states.append(rng.get_state())
        def decode_blocks(blocks):
            app, _ = code.decode_batch(torch.from_numpy(np.asarray([b[1] for b in blocks])).to(dev))
            return ((app < 0.0).cpu().numpy() != np.asarray([b[0] for b in blocks]).astype(bool)).sum(axis=1).astype(float).tolist()

        errs = SD.decode_sharded(decode_blocks, list(zip(xs, ys)), group)
        for j in range(nb):
            nbit += int(errs[j])
            nblockerr += 1 if errs[j] else 0
            nblocks += 1
            ber = nbit / (nblocks * code.N)
            if nblockerr >= MIN_ERRORS or nblocks >= MAX_BLOCKS:
                rng.set_state(states[j + 1])
                break
    return ber


# ----------------------------------------------------------------------------------- Monte-Carlo drivers
def _seqsum(values):
    """Left-to-right fp64 accumulation, as the reference's `BER = BER + ber` loops (sparc_ldpc.py:1229-1233); the
    built-in sum() of CPython >= 3.12 adds Python floats with compensation and would round differently."""
    acc = np.float64(0.0)
    for v in values:
        acc = acc + np.float64(v)
    return acc


def _mc(rng, draw_block, decode_blocks, stop_after, MIN_ERRORS, MAX_BLOCKS, chunk, group=None):
    """Sequential stop-rule replay.  draw_block() consumes the RNG for one block and returns
    its inputs; decode_blocks(list) -> per-block result rows; stop_after(row) says whether the block counts
    as an error block.  Returns the rows the reference would have averaged."""
    rows, nerr = [], 0
    while nerr < MIN_ERRORS and len(rows) < MAX_BLOCKS:
        nb = int(min(chunk, MAX_BLOCKS - len(rows)))
        states, blocks = [], []
        for _ in range(nb):
            states.append(rng.get_state())
            blocks.append(draw_block())
        states.append(rng.get_state())
        out = SD.decode_sharded(decode_blocks, blocks, group)   # (group=None: decode_blocks(blocks))
        for j in range(nb):
            rows.append(out[j])
            nerr += 1 if stop_after(out[j]) else 0
            if nerr >= MIN_ERRORS or len(rows) >= MAX_BLOCKS:
                rng.set_state(states[j + 1])
                return rows
    return rows


def _pair_driver(coded_flow, sp_coded, lp, sp_plain, rng, MIN_ERRORS, MAX_BLOCKS, chunk, seed=0, all_zero=False,
                 group=None, **kw):
    """Blocks of (coded simulation, plain simulation at the same overall rate) as in waterfall() and
    soft_hardinit_plot(): per block the coded draw comes first (sparc_ldpc.py:1218-1231)."""
    su_c = D.make_setup(sp_coded, lp, seed=_matrix_seed(seed, rng))
    su_p = D.make_setup(sp_plain, None)

    def draw_block():
        c = _draw(su_c, 1, sp_coded.sigma, rng, all_zero)
        p = _draw(su_p, 1, sp_plain.sigma, rng)
        return c, p

    def decode_blocks(blocks):
        ic = np.concatenate([b[0][0] for b in blocks]); nc = np.concatenate([b[0][1] for b in blocks])
        ip = np.concatenate([b[1][0] for b in blocks]); npn = np.concatenate([b[1][1] for b in blocks])
        txc, yc = _transmit(su_c, ic, nc)
        stc = coded_flow(su_c, yc, **kw)
        ba, bl = stc.ber(txc, su_c.total_bits)
        txp, yp = _transmit(su_p, ip, npn)
        bp_, _ = D.plain(su_p, yp).ber(txp, su_p.total_bits)
        return [(ba[j], bl[j], bp_[j, 0]) for j in range(len(blocks))]

    return _mc(rng, draw_block, decode_blocks, lambda row: row[2] != 0, MIN_ERRORS, MAX_BLOCKS, chunk, group)


def _maybe_plot(fn):
    try:
        import matplotlib  # noqa: F401
    except Exception:
        print("matplotlib not available: figure skipped (CSV written)")
        return
    fn()


def waterfall(sparcparams, ldpcparams, csv_filename, png_filename, init="soft", pa_param=False, datapoints=10,
              MIN_ERRORS=100, MAX_BLOCKS=500, bpsk=True, sections=512, chunk=64, EbN0_dB=None, rng=None, group=None):
    """BER waterfall (sparc_ldpc.py:1126-1282): same grid, sigma convention (20 log10), hard-coded rate 5/6,
    stop rule and CSV schema.  `EbN0_dB` overrides the default linspace(3, 10, datapoints) grid; returns the
    dict of columns that is written to the CSV.  `group`: decode every chunk's blocks across the ranks of a
    torch.distributed group (every rank must call with the same arguments and an identically seeded `rng`);
    columns and final RNG state equal the one-rank run, rank 0 writes the CSV."""
    rng = _rng(rng)
    L, M = sparcparams.L, sparcparams.M
    logm = np.log2(M)
    p, r_sparc, T = sparcparams.p, sparcparams.r, sparcparams.t
    a, f, C = sparcparams.a, sparcparams.f, sparcparams.C
    nl = logm * sections
    z = int(nl / 24)                                                   # :1153-1155
    ldpcparams = LDPCParams(ldpcparams.standard, ldpcparams.r_ldpc, z)
    n = L * logm / r_sparc
    R = (L * logm - nl * (1 - 5 / 6)) / n                              # :1160 (rate 5/6 hard coded)
    grid = np.linspace(3, 10, datapoints) if EbN0_dB is None else np.asarray(EbN0_dB, dtype=float)
    datapoints = len(grid)
    cols = {k: np.zeros(datapoints) for k in ("BER_amp_1", "BER_ldpc", "BER_amp_2", "BER_ldpc_2", "BER_plain", "BER_bpsk")}
    flows = {"soft": (D.soft, dict(soft_iter=2)), "hard": (D.hard_init, {}), "originalHard": (D.original_hard, {})}
    if init not in flows:
        raise ValueError("%s is not a valid initialisation. Please change to 'soft', 'hard' or 'originalHard'" % init)
    flow, kw = flows[init]
    for i, ebno_db in enumerate(grid):
        ebno = 10 ** (ebno_db / 20)
        if bpsk:
            cols["BER_bpsk"][i] = sim_ldpc(ldpcparams, np.sqrt((1 / ebno) / 2), MIN_ERRORS, MAX_BLOCKS, rng=rng, group=group)  # :1189-1193
        snr = ebno / (1 / (2 * R))
        sigma = np.sqrt(p / snr)
        C = 0.5 * np.log2(1 + p / (sigma ** 2))
        if pa_param and a is None:
            a = f = r_sparc / C                                        # frozen from the first point (:1202-1204)
        sp_c = SPARCParams(L, M, sigma, p, r_sparc, T, a, f, C)
        sp_p = SPARCParams(L, M, sigma, p, R, T, a, f, C)
        rows = _pair_driver(flow, sp_c, ldpcparams, sp_p, rng, MIN_ERRORS, MAX_BLOCKS, chunk, group=group, **kw)
        nb = len(rows)
        amp = np.array([np.pad(r_[0], (0, max(0, 2 - len(r_[0])))) for r_ in rows])
        ldp = np.array([np.pad(r_[1], (0, max(0, 2 - len(r_[1])))) for r_ in rows])  # hard/originalHard: 0 appended (:1223-1227)
        cols["BER_amp_1"][i], cols["BER_amp_2"][i] = amp[:, 0].sum() / nb, amp[:, 1].sum() / nb
        cols["BER_ldpc"][i], cols["BER_ldpc_2"][i] = ldp[:, 0].sum() / nb, ldp[:, 1].sum() / nb
        cols["BER_plain"][i] = _seqsum(r_[2] for r_ in rows) / nb
    fields = ["EbN0_dB", "BER_amp_1", "BER_ldpc", "BER_amp_2", "BER_ldpc_2", "BER_plain", "BER_bpsk"]  # :1260
    if SD.group_info(group)[0] == 0:
        with open(csv_filename, "a") as fh:
            w = csv.DictWriter(fh, fieldnames=fields)
            w.writeheader()
            for k in range(datapoints):
                w.writerow(dict(EbN0_dB=grid[k], **{c: cols[c][k] for c in fields[1:]}))
    cols["EbN0_dB"] = grid
    return cols


def soft_hardinit_plot(sparcparams, ldpcparams, csv_filename, png_filename, sections, datapoints=10, MIN_ERRORS=100,
                       MAX_BLOCKS=500, soft_iter=3, threshold=0.6, chunk=64, SIGMA=None, rng=None, group=None):
    """Threshold-initialised exchange sweep (sparc_ldpc.py:1435-1581), sigma grid linspace(0.9, 1.4).  `group`: as in
    waterfall()."""
    rng = _rng(rng)
    L, M = sparcparams.L, sparcparams.M
    logm = np.log2(M)
    p, r_sparc, T = sparcparams.p, sparcparams.r, sparcparams.t
    standard, r_ldpc, z = ldpcparams.standard, ldpcparams.r_ldpc, ldpcparams.z
    Rldpc = {"5/6": 5 / 6, "1/2": 1 / 2, "0.45": 0.45, "3/8": 3 / 8}[r_ldpc]
    nl = logm * sections
    ieee = standard in ("802.11n", "802.16")
    if z is None:
        z = int(nl / 24) if ieee else int(nl / 40)                     # :1470-1476
    ldpcparams = LDPCParams(standard, r_ldpc, z)
    n = L * logm / r_sparc
    R = (L * logm - nl * (1 - Rldpc)) / n
    SIGMA = np.linspace(0.9, 1.4, datapoints) if SIGMA is None else np.asarray(SIGMA, dtype=float)
    datapoints = len(SIGMA)
    BER_amp, BER_ldpc = np.zeros((datapoints, soft_iter)), np.zeros((datapoints, soft_iter))
    BER_plain = np.zeros(datapoints)
    for i, sigma in enumerate(SIGMA):
        rows = _pair_driver(D.threshold, SPARCParams(L, M, sigma, p, r_sparc, T), ldpcparams,
                            SPARCParams(L, M, sigma, p, R, T), rng, MIN_ERRORS, MAX_BLOCKS, chunk,
                            seed=0 if ieee else None, all_zero=not ieee, group=group, soft_iter=soft_iter, thr=threshold)
        nb = len(rows)
        BER_amp[i] = np.sum([r_[0] for r_ in rows], axis=0) / nb
        BER_ldpc[i] = np.sum([r_[1] for r_ in rows], axis=0) / nb
        BER_plain[i] = _seqsum(r_[2] for r_ in rows) / nb
    EbN0_dB = 20 * np.log10(1 / (2 * R) * (p / SIGMA ** 2))            # :1531-1533
    if SD.group_info(group)[0] == 0:
        with open(csv_filename, "a") as fh:
            w = csv.DictWriter(fh, fieldnames=["EbN0_dB", "BER_amp", "BER_ldpc", "BER_plain"])  # :1537
            w.writeheader()
            for k in range(datapoints):
                w.writerow({"EbN0_dB": EbN0_dB[k], "BER_amp": BER_amp[k, :], "BER_ldpc": BER_ldpc[k, :], "BER_plain": BER_plain[k]})
    return dict(EbN0_dB=EbN0_dB, BER_amp=BER_amp, BER_ldpc=BER_ldpc, BER_plain=BER_plain)


def soft_hard_plot(soft, hard, sec, soft_iter, sparcparams, ldpcparams, csv_filename, png_filename, datapoints=10,
                   MIN_ERRORS=100, MAX_BLOCKS=500, chunk=64, SIGMA=None, rng=None, group=None):
    """Soft vs original-hard exchange sweep (sparc_ldpc.py:1285-1432), sigma grid linspace(0.8, 0.4).  `group`: as in
    waterfall()."""
    rng = _rng(rng)
    L, M = sparcparams.L, sparcparams.M
    logm = np.log2(M)
    p, r_sparc, T = sparcparams.p, sparcparams.r, sparcparams.t
    a, f, C = sparcparams.a, sparcparams.f, sparcparams.C
    nl = logm * sec
    z = int(nl / 24)
    ldpcparams = LDPCParams(ldpcparams.standard, ldpcparams.r_ldpc, z)
    n = L * logm / r_sparc
    R = (L * logm - nl * (1 - 5 / 6)) / n
    SIGMA = np.linspace(0.8, 0.4, datapoints) if SIGMA is None else np.asarray(SIGMA, dtype=float)
    datapoints = len(SIGMA)
    out = dict(BER_sparc=np.zeros(datapoints))
    if soft:
        out["BER_ldpc_soft"], out["BER_amp_soft"] = np.zeros((datapoints, soft_iter)), np.zeros((datapoints, soft_iter + 1))
    if hard:
        out["BER_amp_hard"], out["BER_ldpc_hard"] = np.zeros((datapoints, 2)), np.zeros(datapoints)

    def single(flow, sp, lp, stop, **kw):
        su = D.make_setup(sp, lp)

        def decode_blocks(blocks):
            tx, y = _transmit(su, np.concatenate([b[0] for b in blocks]), np.concatenate([b[1] for b in blocks]))
            ba, bl = flow(su, y, **kw).ber(tx, su.total_bits)
            return [(ba[j], bl[j]) for j in range(len(blocks))]

        return _mc(rng, lambda: _draw(su, 1, sp.sigma, rng), decode_blocks, stop, MIN_ERRORS, MAX_BLOCKS, chunk, group)

    for i, sigma in enumerate(SIGMA):
        sp_c = SPARCParams(L, M, sigma, p, r_sparc, T, a, f, C)
        sp_p = SPARCParams(L, M, sigma, p, R, T, a, f, C)
        # plain SPARC: exactly MIN_ERRORS runs (:1341-1344)
        su_p = D.make_setup(sp_p, None)
        blocks = [_draw(su_p, 1, sigma, rng) for _ in range(MIN_ERRORS)]

        def decode_plain(bl):
            tx, y = _transmit(su_p, np.concatenate([b[0] for b in bl]), np.concatenate([b[1] for b in bl]))
            return D.plain(su_p, y).ber(tx, su_p.total_bits)[0][:, 0].tolist()

        out["BER_sparc"][i] = np.sum(SD.decode_sharded(decode_plain, blocks, group)) / MIN_ERRORS
        if soft:
            rows = single(D.soft, sp_c, ldpcparams, lambda row: row[1][0] != 0, soft_iter=soft_iter)  # :1359
            out["BER_amp_soft"][i] = np.sum([r_[0] for r_ in rows], axis=0) / len(rows)
            out["BER_ldpc_soft"][i] = np.sum([r_[1] for r_ in rows], axis=0) / len(rows)
        if hard:
            rows = single(D.original_hard, sp_c, ldpcparams, lambda row: row[1][0] != 0)                # :1381
            out["BER_amp_hard"][i] = np.sum([np.pad(r_[0], (0, 2 - len(r_[0]))) for r_ in rows], axis=0) / len(rows)
            out["BER_ldpc_hard"][i] = np.sum([r_[1][0] for r_ in rows]) / len(rows)
    EbN0_dB = 20 * np.log10(1 / (2 * R) * (p / SIGMA ** 2))
    with open(csv_filename if SD.group_info(group)[0] == 0 else os.devnull, "a") as fh:
        if soft:
            w = csv.DictWriter(fh, fieldnames=["EbN0_dB", "BER_sparc", "BER_ldpc_soft", "BER_amp_soft"])  # :1400
            w.writeheader()
            for k in range(datapoints):
                w.writerow({"EbN0_dB": EbN0_dB[k], "BER_sparc": out["BER_sparc"][k],
                            "BER_ldpc_soft": out["BER_ldpc_soft"][k, :], "BER_amp_soft": out["BER_amp_soft"][k, :]})
        if hard:
            w = csv.DictWriter(fh, fieldnames=["EbN0_dB", "BER_sparc", "BER_ldpc_hard", "BER_amp_hard"])  # :1406
            w.writeheader()
            for k in range(datapoints):
                w.writerow({"EbN0_dB": EbN0_dB[k], "BER_sparc": out["BER_sparc"][k],
                            "BER_ldpc_hard": out["BER_ldpc_hard"][k], "BER_amp_hard": out["BER_amp_hard"][k, :]})
    out["EbN0_dB"] = EbN0_dB
    return out
