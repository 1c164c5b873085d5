"""CPU oracle: numpy/C restatement of the reference's SPARC-AMP + LDPC hot path.

TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's
`cpu_baseline` / `--impl reference` legs may import this module; nothing under
sparc_ldpc_b200/ does.  Every function cites the reference lines it restates
(paths relative to /root/reference).

Parity status: PINNED.  tests/test_oracle_cpu.py checks this module against
golden vectors produced by the UNMODIFIED reference run through
oracle/ref_harness.py (generator: tests/golden/gen_golden.py) and against the
reference's own known answers (removed.py:40-49, :203-204,
ldpc802.16.81.h arrays, test_ldpc.py properties).  The numpy calls below are the
same numpy calls the reference makes wherever summation order matters
(np.sum -> pairwise, np.cumsum -> sequential), so on one machine the oracle
reproduces the reference bit for bit.

The heavy inner pieces (w-point Walsh-Hadamard transform, BP) live in
oracle/oracle.c, built by oracle/Makefile into oracle/_build/liboracle.so.
"""
import ctypes
import json
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
MAX_ITCOUNT = 200  # ldpc/src/c_ldpc.c:7


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "_build", "liboracle.so")
        if not os.path.isfile(path):
            raise RuntimeError("oracle/_build/liboracle.so missing: run `make -C oracle`")
        L = ctypes.CDLL(path)
        vp, cl, ci, cd = ctypes.c_void_p, ctypes.c_long, ctypes.c_int, ctypes.c_double
        L.orc_fht_inplace.argtypes = [vp, cl]
        L.orc_fht_inplace.restype = None
        for f in (L.orc_block_Ax, L.orc_block_Ay):
            f.argtypes = [vp, cl, cl, cl, cl, vp, vp]
            f.restype = ci
        L.orc_Lxor.argtypes = [cd, cd, ci]
        L.orc_Lxor.restype = cd
        L.orc_Lxfb.argtypes = [vp, cl, ci]
        L.orc_Lxfb.restype = cd
        for f in (L.orc_sumprod2, L.orc_sumprod):
            f.argtypes = [vp, vp, vp, vp, ci, ci, ci, vp, ci]
            f.restype = ci
        L.orc_minsum.argtypes = [vp, vp, vp, vp, ci, ci, ci, vp, cd, ci]
        L.orc_minsum.restype = ci
        _LIB = L
    return _LIB


# --------------------------------------------------------------------------- operators
def fht_inplace(x):
    """ldpc/sparc_ldpc.py:19-29 (in-repo definition of pyfht.fht_inplace)."""
    assert x.dtype == np.float64 and x.flags.c_contiguous
    lib().orc_fht_inplace(x.ctypes.data, x.size)


def transform_width(M, n):
    """ldpc/sparc_ldpc.py:54,110: w = 2**ceil(log2(max(m+1, n+1)))."""
    return 2 ** int(np.ceil(np.log2(max(M + 1, n + 1))))


def make_ordering(L, M, n, seed=0):
    """ldpc/sparc_ldpc.py:110-117: cumulative in-place shuffles of arange(1, w)."""
    w = transform_width(M, n)
    rng = np.random.RandomState(seed)
    ordering = np.empty((L, n), dtype=np.uint32)
    idxs = np.arange(1, w, dtype=np.uint32)
    for ll in range(L):
        rng.shuffle(idxs)
        ordering[ll] = idxs[:n]
    return ordering


def _operators(L, M, n, ordering):
    ordering = np.ascontiguousarray(ordering[:L], dtype=np.uint32)
    assert ordering.shape == (L, n)
    w = transform_width(M, n)
    rt_n = np.sqrt(n)

    def Ab(b):  # sparc_ldpc.py:143-144 over :120-126, :65-70
        b = np.ascontiguousarray(np.asarray(b, dtype=np.float64).reshape(-1))
        assert b.size == L * M
        out = np.empty(n)
        lib().orc_block_Ax(ordering.ctypes.data, L, n, M, w, b.ctypes.data, out.ctypes.data)
        return out.reshape(-1, 1) / rt_n

    def Az(z):  # sparc_ldpc.py:145-146 over :128-134, :72-77
        z = np.ascontiguousarray(np.asarray(z, dtype=np.float64).reshape(-1))
        assert z.size == n
        out = np.empty(L * M)
        lib().orc_block_Ay(ordering.ctypes.data, L, n, M, w, z.ctypes.data, out.ctypes.data)
        return out.reshape(-1, 1) / rt_n

    return Ab, Az


def sparc_transforms(L, M, n, seed=0):
    """ldpc/sparc_ldpc.py:140-147."""
    ordering = make_ordering(L, M, n, seed)
    Ab, Az = _operators(L, M, n, ordering)
    return Ab, Az, ordering


def sparc_transforms_shorter(L, M, n, ordering):
    """ldpc/sparc_ldpc.py:154-168: the same operator on the first L rows of `ordering`."""
    return _operators(L, M, n, ordering)


def pa_parameterised(L, C, P, a, f):
    """ldpc/sparc_ldpc.py:172-186 (IndexError when f >= 1, as in the reference)."""
    pa = 2 ** (-2 * a * C * np.arange(L) / L)
    pa[int(f * L):] = pa[int(f * L)]
    pa /= pa.sum() / P
    return pa


# --------------------------------------------------------------------------- AMP
def amp(y, Pl, L, M, T, Ab, Az, beta0=None, trace=None):
    """ldpc/sparc_ldpc.py:189-222 (and amp_test.py:14-50 for the returned t).

    Returns (beta (L*M,1), t) where t is what amp_test returns: the iteration
    index at which tau == last_tau fired, or T-1 when all T iterations ran.
    `trace`, if a list, receives (tau2, beta_after_update) per executed iteration.
    """
    P = np.sum(Pl)
    n = y.size
    if beta0 is None:
        beta = np.zeros((L * M, 1))
        z = y
    else:
        beta = np.asarray(beta0, dtype=np.float64).reshape(L * M, 1)
        z = y - Ab(beta)
    last_tau = 0
    t = 0
    for t in range(T):
        tau = np.sqrt(np.sum(z ** 2) / n)
        if tau == last_tau:
            return beta, t
        last_tau = tau
        s = beta + Az(z)
        rt_n_Pl = np.sqrt(n * Pl).repeat(M).reshape(-1, 1)
        u = s * rt_n_Pl / tau ** 2
        max_u = u.max()
        exps = np.exp(u - max_u)
        sums = exps.reshape(L, M).sum(axis=1).repeat(M).reshape(-1, 1)
        beta = (rt_n_Pl * exps / sums).reshape(-1, 1)
        z = y - Ab(beta) + (z / tau ** 2) * (P - np.sum(beta ** 2) / n)
        if trace is not None:
            trace.append((float(tau ** 2), beta.reshape(-1).copy()))
    return beta, t


# --------------------------------------------------------------------------- section <-> bit maps
def sp2bp(beta, L, M):
    """ldpc/sparc_ldpc.py:257-281: P(bit=1), MSB first, accumulated over ascending j
    (sequential adds; np.cumsum is strictly sequential, so this is bit-identical to
    the reference's Python loop)."""
    logm = int(np.log2(M))
    b = np.asarray(beta, dtype=np.float64).reshape(L, M)
    p = np.zeros((L, logm))
    j = np.arange(M)
    for logi in range(logm):
        sel = j[(j >> logi) & 1 == 1]
        p[:, logm - logi - 1] = np.cumsum(b[:, sel], axis=1)[:, -1]
    return p.reshape(-1)


def sp2bp_loops(beta, L, M):
    """Literal triple loop of ldpc/sparc_ldpc.py:268-281 -- small cases only."""
    logm = int(np.log2(M))
    p = np.zeros(logm * L)
    for a in range(L):
        bl = beta[a * M:(a + 1) * M]
        for logi in range(logm):
            bpos = (a + 1) * logm - logi - 1
            i = 2 ** logi
            k = i
            while k < M:
                for j in range(k, k + i):
                    p[bpos] = p[bpos] + bl[j]
                k += 2 * i
    return p


def bitwise_to_llr(p):
    """ldpc/sparc_ldpc.py:667-669 (also :477-479, :823-825, :979-981, :1026-1028,
    amp_exit.py:246-248): log(1-p) - log(p), NaN -> 0, +-inf -> +-DBL_MAX."""
    with np.errstate(divide="ignore", invalid="ignore"):
        llr = np.log(1 - p) - np.log(p)
    return np.nan_to_num(llr)


def bp2sp(v, L, M):
    """ldpc/sparc_ldpc.py:283-314: product over the logM bits (MSB first, sequential
    np.prod order), then divide each section by its np.sum (pairwise)."""
    logm = int(np.log2(M))
    v = np.asarray(v, dtype=np.float64).reshape(L, logm)
    m = np.arange(M)
    sp = np.ones((L, M))
    for jb in range(logm):
        bit = ((m >> (logm - 1 - jb)) & 1).astype(bool)
        one = v[:, jb:jb + 1] * np.ones((1, M))      # bp**1 * (1-bp)**0
        zero = np.ones((1, M)) * (1 - v[:, jb:jb + 1])  # bp**0 * (1-bp)**1
        sp = sp * np.where(bit[None, :], one, zero)
    sp = sp / np.sum(sp, axis=1, keepdims=True)
    return sp.reshape(-1)


def bits2indices(bits, M):
    """ldpc/sparc_ldpc.py:317-341 (MSB first)."""
    logm = int(np.log2(M))
    b = np.asarray(bits).astype(np.int64).reshape(-1, logm)
    wts = 1 << np.arange(logm - 1, -1, -1)
    return (b * wts).sum(axis=1).tolist()


def count_bit_errors(idx_a, idx_b):
    """sum(bin(a^b).count('1')) of ldpc/sparc_ldpc.py:462,500,650,..."""
    return sum(bin(int(a) ^ int(b)).count("1") for a, b in zip(idx_a, idx_b))


def ber_from_LLRs(M, LLR, input_indices, total_bits):
    """ldpc/sparc_ldpc.py:343-356."""
    return count_bit_errors(input_indices, bits2indices(LLR < 0.0, M)) / total_bits


def argmax_sections(beta, L, M):
    """ldpc/sparc_ldpc.py:640-643: first maximum wins (np.argmax)."""
    return np.argmax(np.asarray(beta).reshape(L, M), axis=1).tolist()


# --------------------------------------------------------------------------- LDPC code object
_PROTO_DB = None


def _proto_db():
    global _PROTO_DB
    if _PROTO_DB is None:
        path = os.path.join(os.path.dirname(_HERE), "sparc_ldpc_b200", "data", "protographs.json")
        with open(path) as f:
            _PROTO_DB = json.load(f)
    return _PROTO_DB


def load_proto(standard, rate, z, ptype="A"):
    """Protograph tables of ldpc/py/ldpc.py:59-660, read from the extracted data file."""
    db = _proto_db()
    if standard == "802.11n":
        if z not in (27, 54, 81):
            raise NameError("802.11n invalid z (must be 27,54 or 81)")
        key = "|".join([standard, rate, "-", str(z)])
    elif standard == "802.16" and rate in ("2/3", "3/4"):
        if ptype not in ("A", "B"):
            raise NameError("802.16 type must be either A or B")
        key = "|".join([standard, rate, ptype, "*"])
    else:
        key = "|".join([standard, rate, "-", "*"])
    if key not in db:
        raise NameError("unknown LDPC code %s" % key)
    e = db[key]
    proto = -np.ones((e["rows"], e["cols"]), dtype=np.int64)
    for r, c, s in e["edges"]:
        proto[r, c] = s
    return proto


class Code:
    """ldpc/py/ldpc.py:6-22 + prepare_decoder :694-786 + encode :790-850 + decode :855-930."""

    def __init__(self, standard="802.11n", rate="1/2", z=27, ptype="A"):
        self.standard, self.rate, self.z, self.ptype = standard, rate, z, ptype
        self.proto = load_proto(standard, rate, z, ptype)
        self.vdeg, self.cdeg, self.intrlv = self.prepare_decoder()
        self.Nv, self.Nc, self.Nmsg = len(self.vdeg), len(self.cdeg), len(self.intrlv)
        self.N = self.Nv
        self.K = self.Nv - self.Nc

    def prepare_decoder(self):
        """ldpc.py:694-786.  The reference scans for the first unused port of the
        check and of the variable for every edge while walking the protograph in
        row-major order; a per-node fill counter gives the same assignment."""
        proto, z = self.proto, self.z
        cdeg = np.repeat(np.sum(proto != -1, 1), z)
        vdeg = np.repeat(np.sum(proto != -1, 0), z)
        cum_c = np.insert(np.cumsum(cdeg), 0, 0)
        cum_v = np.insert(np.cumsum(vdeg), 0, 0)
        c_fill = np.zeros(len(cdeg), dtype=np.int64)
        v_fill = np.zeros(len(vdeg), dtype=np.int64)
        pre = -np.ones(int(np.sum(cdeg)), dtype=np.int64)
        xp, yp = np.nonzero(proto != -1)
        k = np.arange(z)
        for j in range(xp.size):
            off = proto[xp[j], yp[j]]
            cind = xp[j] * z + k
            vind = yp[j] * z + (k + off) % z
            xi = cum_c[cind] + c_fill[cind]
            yi = cum_v[vind] + v_fill[vind]
            c_fill[cind] += 1
            v_fill[vind] += 1
            pre[xi] = yi
        return vdeg.astype(np.int64), cdeg.astype(np.int64), np.argsort(pre).astype(np.int64)

    def pcmat(self):
        """ldpc.py:666-691."""
        proto, z = self.proto, self.z
        H = np.zeros((z * proto.shape[0], z * proto.shape[1]), dtype=int)
        for r, c in zip(*np.nonzero(proto != -1)):
            H[r * z:(r + 1) * z, c * z:(c + 1) * z] = np.roll(np.eye(z, dtype=int), proto[r, c] % z, 1)
        return H

    def encode(self, info):
        """ldpc.py:790-850: systematic QC encoding with dual-diagonal back-substitution."""
        z, proto = self.z, self.proto
        Mp, Np = proto.shape
        Kp = Np - Mp
        if len(info) != Kp * z:
            raise NameError("information word length not compatible with proto and z")
        x = np.zeros((Np, z), dtype=int)
        x[:Kp] = np.asarray(info, dtype=int).reshape(Kp, z)
        p = np.zeros((Mp, z), dtype=int)
        for j in range(Mp):
            for k in np.nonzero(proto[j, :Kp] != -1)[0]:
                p[j] += np.roll(x[k], -proto[j, k])
        p %= 2
        tp = p.sum(0) % 2
        toff = np.zeros(z, dtype=int)
        for j in np.nonzero(proto[:, Kp] != -1)[0]:
            toff[proto[j, Kp] % z] += 1
        nz = np.nonzero(toff % 2)[0]
        if len(nz) != 1:
            raise NameError("The offsets in colum Kp+1 of proto do not add to a single offset")
        x[Kp] = np.roll(tp, nz[0])
        for j in range(Mp - 1):
            cur = Kp + j + 1
            acc = p[j].copy()
            for k in np.nonzero(proto[j, Kp:cur] != -1)[0]:
                acc += np.roll(x[Kp + k], -proto[j, Kp + k])
            x[cur] = acc % 2
        return x.reshape(-1)

    def decode(self, ch, dectype="sumprod2", corr_factor=0.7, max_it=MAX_ITCOUNT):
        """ldpc.py:855-930 over c_ldpc.c (restated in oracle.c)."""
        ch = np.ascontiguousarray(ch, dtype=np.float64)
        if len(ch) != self.Nv:
            raise NameError("Channel inputs not consistent with variable degrees")
        app = np.zeros(self.Nv)
        args = (ch.ctypes.data, self.vdeg.ctypes.data, self.cdeg.ctypes.data, self.intrlv.ctypes.data,
                self.Nv, self.Nc, self.Nmsg, app.ctypes.data)
        if dectype == "sumprod2":
            it = lib().orc_sumprod2(*args, max_it)
        elif dectype == "sumprod":
            it = lib().orc_sumprod(*args, max_it)
        elif dectype == "minsum":
            it = lib().orc_minsum(*args, corr_factor, max_it)
        else:
            raise NameError("Decoder type unknonwn")
        return app, it


def Lxor(L1, L2, corr=1):
    return lib().orc_Lxor(L1, L2, corr)


def Lxfb(Lin, corr=1):
    a = np.array(Lin, dtype=np.float64)
    tot = lib().orc_Lxfb(a.ctypes.data, len(a), corr)
    return tot, a


_CODE_CACHE = {}


def get_code(standard, rate, z, ptype="A"):
    key = (standard, rate, z, ptype)
    if key not in _CODE_CACHE:
        _CODE_CACHE[key] = Code(standard, rate, z, ptype)
    return _CODE_CACHE[key]


# --------------------------------------------------------------------------- link simulations
class SPARCParams:
    """ldpc/sparc_ldpc.py:227-246."""

    def __init__(self, L, M, sigma, p, r, t, a=None, f=None, C=None):
        self.L, self.M, self.sigma, self.p, self.r, self.t = L, M, sigma, p, r, t
        self.a, self.f, self.C = a, f, C


class LDPCParams:
    """ldpc/sparc_ldpc.py:250-255."""

    def __init__(self, standard, r_ldpc, z, ptype="A"):
        self.standard, self.r_ldpc, self.z, self.ptype = standard, r_ldpc, z, ptype


def _setup(sp):
    L, M = sp.L, sp.M
    n = int(L * np.log2(M) / sp.r)
    logm = int(np.log2(M))
    Pl = sp.p / L * np.ones(L) if sp.a is None else pa_parameterised(L, sp.C, sp.p, sp.a, sp.f)
    return L, M, n, logm, int(logm * L), Pl


def _draw_message(code, total_bits, rng):
    """RNG draw order of sparc_ldpc.py:419-426 / :606-612 / :767-773 / :919-926."""
    if code is None:
        return rng.randint(0, 2, total_bits).tolist(), 0, 0
    protected = rng.randint(0, 2, code.K).tolist()
    ldpc_bits = code.encode(protected).tolist()
    unprotected = rng.randint(0, 2, int(total_bits - code.N)).tolist()
    return unprotected + ldpc_bits, code.N, code.K


def _encode_and_channel(idx, L, M, n, Pl, sigma, rng, seed=0):
    """sparc_ldpc.py:433-446: beta_0 one-hot * sqrt(n P_l); x = A beta_0; y = x + sigma * randn(n,1)."""
    Ab, Az, ordering = sparc_transforms(L, M, n, seed)
    b0 = np.zeros((L * M, 1))
    for l in range(L):
        b0[l * M + idx[l]] = np.sqrt(n * Pl[l])
    x = Ab(b0)
    noise = rng.randn(n, 1) * sigma
    return (x + noise).reshape(-1, 1), Ab, Az, ordering


def amp_ldpc_sim(sp, lp=None, rng=np.random, record=None):
    """ldpc/sparc_ldpc.py:359-545 ("original hard" exchange; plain SPARC when lp is None)."""
    L, M, n, logm, total_bits, Pl = _setup(sp)
    code = None if lp is None else get_code(lp.standard, lp.r_ldpc, lp.z, lp.ptype)
    if code is not None:
        assert code.N <= L * logm and code.N % logm == 0
    bits, nl, kl = _draw_message(code, total_bits, rng)
    idx = bits2indices(bits, M)
    y, Ab, Az, ordering = _encode_and_channel(idx, L, M, n, Pl, sp.sigma, rng)
    beta, _ = amp(y, Pl, L, M, sp.t, Ab, Az)
    beta = beta.reshape(-1)
    rx = argmax_sections(beta, L, M)
    ber_amp = count_bit_errors(idx, rx) / total_bits
    ber_ldpc = ber_ldpc_amp = None
    if record is not None:
        record.update(idx=idx, y=y.reshape(-1).copy(), beta1=beta.copy())
    if code is not None:
        post = beta / np.sqrt(n * np.repeat(Pl, M))
        ls = int(nl / logm)
        llr = bitwise_to_llr(sp2bp(post[(L - ls) * M:], ls, M))
        app, it = code.decode(llr)
        out_l = bits2indices(app < 0.0, M)
        rx[L - ls:] = out_l
        ber_ldpc = count_bit_errors(idx, rx) / total_bits
        if record is not None:
            record.update(llr1=llr.copy(), app1=app.copy(), it1=it)
        if L - ls > 0:
            bl = np.zeros((L * M, 1))
            for i, l in enumerate(range(L - ls, L)):
                bl[l * M + out_l[i]] = np.sqrt(n * Pl[l])
            y_new = y - Ab(bl)
            Lu = L - ls
            Ab2, Az2 = sparc_transforms_shorter(Lu, M, n, ordering)
            b2, _ = amp(y_new, Pl[:Lu], Lu, M, sp.t, Ab2, Az2)
            rx[:Lu] = argmax_sections(b2.reshape(-1), Lu, M)
            ber_ldpc_amp = count_bit_errors(idx, rx) / total_bits
            if record is not None:
                record.update(beta2=b2.reshape(-1).copy())
    R = (L * logm - (nl - kl)) / n
    return ber_amp, ber_ldpc, ber_ldpc_amp, R


def soft_amp_ldpc_sim(sp, lp, soft_iter, rng=np.random, record=None):
    """ldpc/sparc_ldpc.py:547-712 (soft exchange)."""
    L, M, n, logm, total_bits, Pl = _setup(sp)
    code = get_code(lp.standard, lp.r_ldpc, lp.z, lp.ptype)
    assert code.N <= L * logm and code.N % logm == 0
    bits, nl, kl = _draw_message(code, total_bits, rng)
    idx = bits2indices(bits, M)
    y, Ab, Az, _ = _encode_and_channel(idx, L, M, n, Pl, sp.sigma, rng)
    beta, _ = amp(y, Pl, L, M, sp.t, Ab, Az)
    beta = beta.reshape(-1)
    rx = argmax_sections(beta, L, M)
    ber_amp = [count_bit_errors(idx, rx) / total_bits]
    ber_ldpc = []
    scale = np.sqrt(n * np.repeat(Pl, M))
    ls = int(nl / logm)
    if record is not None:
        record.update(idx=idx, y=y.reshape(-1).copy(), beta=[beta.copy()], llr=[], app=[], it=[])
    for _ in range(soft_iter):
        post = beta / scale
        llr = bitwise_to_llr(sp2bp(post[(L - ls) * M:], ls, M))
        app, it = code.decode(llr)
        rx[L - ls:] = bits2indices(app < 0.0, M)
        ber_ldpc.append(count_bit_errors(idx, rx) / total_bits)
        with np.errstate(over="ignore"):
            bw = 1 / (1 + np.exp(app))
        post[M * (L - ls):] = bp2sp(bw, ls, M)
        beta, _ = amp(y, Pl, L, M, sp.t, Ab, Az, post * scale)
        beta = beta.reshape(-1)
        rx = argmax_sections(beta, L, M)
        ber_amp.append(count_bit_errors(idx, rx) / total_bits)
        if record is not None:
            record["llr"].append(llr.copy()); record["app"].append(app.copy())
            record["it"].append(it); record["beta"].append(beta.copy())
    return ber_amp, ber_ldpc, (L * logm - (nl - kl)) / n


def hardinitbeta_amp_ldpc_sim(sp, lp, rng=np.random, record=None):
    """ldpc/sparc_ldpc.py:715-860 (hard-decided beta as the AMP initialisation)."""
    L, M, n, logm, total_bits, Pl = _setup(sp)
    code = get_code(lp.standard, lp.r_ldpc, lp.z, lp.ptype)
    assert code.N <= L * logm and code.N % logm == 0
    bits, nl, kl = _draw_message(code, total_bits, rng)
    idx = bits2indices(bits, M)
    y, Ab, Az, _ = _encode_and_channel(idx, L, M, n, Pl, sp.sigma, rng)
    beta, _ = amp(y, Pl, L, M, sp.t, Ab, Az)
    beta = beta.reshape(-1)
    rx = argmax_sections(beta, L, M)
    ber_amp = [count_bit_errors(idx, rx) / total_bits]
    post = beta / np.sqrt(n * np.repeat(Pl, M))
    ls = int(nl / logm)
    llr = bitwise_to_llr(sp2bp(post[(L - ls) * M:], ls, M))
    app, it = code.decode(llr)
    rx[L - ls:] = bits2indices(app < 0.0, M)
    ber_ldpc = [count_bit_errors(idx, rx) / total_bits]
    bl = np.zeros((L * M, 1))
    for l in range(L):
        bl[l * M + rx[l]] = np.sqrt(n * Pl[l])
    beta2, _ = amp(y, Pl, L, M, sp.t, Ab, Az, bl)
    rx2 = argmax_sections(beta2.reshape(-1), L, M)
    ber_amp.append(count_bit_errors(idx, rx2) / total_bits)
    if record is not None:
        record.update(idx=idx, y=y.reshape(-1).copy(), beta1=beta.copy(), llr1=llr.copy(), app1=app.copy(),
                      it1=it, beta2=beta2.reshape(-1).copy())
    return ber_amp, ber_ldpc, (L * logm - (nl - kl)) / n


def hard_initialisation(beta, L, M, n, ordering, y, Pl, Ab, threshold=0.5, ldpc_sections=None):
    """ldpc/amp_exit.py:56-122.  Mutates `beta` in place like the reference (:79, :96-103)."""
    if ldpc_sections is None:
        ldpc_sections = L
    amp_sections = []
    for l in range(L):
        sec = beta[l * M:(l + 1) * M]
        hits = np.where(sec > threshold)[0] if l >= L - ldpc_sections else np.array([], dtype=int)
        if hits.size == 1:
            sec[:] = 0
            sec[hits[0]] = np.sqrt(n * Pl[l])
        else:
            sec[:] = 0
            amp_sections.append(l)
    y_new = y - Ab(beta)
    La = len(amp_sections)
    if La > 0:
        Ab_new, Az_new = sparc_transforms_shorter(La, M, n, ordering[amp_sections, :])
    else:
        Ab_new = Az_new = None
    return y_new, Ab_new, Az_new, amp_sections, La


def soft_amp_ldpc_hardinit(sp, lp, soft_iter, threshold, rng=np.random, record=None):
    """ldpc/sparc_ldpc.py:862-1046 (threshold-initialised exchange)."""
    L, M, n, logm, total_bits, Pl = _setup(sp)
    code = get_code(lp.standard, lp.r_ldpc, lp.z, lp.ptype)
    nl, kl = code.N, code.K
    assert nl <= total_bits and nl % logm == 0
    if lp.standard in ("802.11n", "802.16"):
        bits, _, _ = _draw_message(code, total_bits, rng)
        seed = 0
    else:
        bits = np.zeros(total_bits)
        seed = None
    idx = bits2indices(bits, M)
    y, Ab, Az, ordering = _encode_and_channel(idx, L, M, n, Pl, sp.sigma, rng, seed)
    beta, _ = amp(y, Pl, L, M, sp.t, Ab, Az)
    beta = beta.reshape(-1)
    rx = argmax_sections(beta, L, M)
    ber_amp = [count_bit_errors(idx, rx) / total_bits]
    ber_ldpc = []
    post = beta / np.sqrt(n * np.repeat(Pl, M))
    ls = int(nl / logm)
    LLR = bitwise_to_llr(sp2bp(post, L, M))
    if record is not None:
        record.update(idx=idx, y=y.reshape(-1).copy(), beta1=beta.copy(), llr0=LLR.copy(), stages=[])
    for i in range(soft_iter):
        app, it = code.decode(LLR[(L - ls) * logm:])
        LLR[(L - ls) * logm:] = app
        ber_ldpc.append(ber_from_LLRs(M, LLR, idx, total_bits))
        if i == soft_iter - 1:
            break
        with np.errstate(over="ignore"):
            bw = 1 / (1 + np.exp(app))
        post[M * (L - ls):] = bp2sp(bw, ls, M)
        y_new, Ab_n, Az_n, amp_sections, La = hard_initialisation(post, L, M, n, ordering, y, Pl, Ab, threshold, ls)
        if La > 0:
            bT, _ = amp(y_new, Pl[amp_sections], La, M, sp.t, Ab_n, Az_n)
            sec = bT.reshape(-1) / np.sqrt(n * np.repeat(Pl[amp_sections], M))
            llr_a = bitwise_to_llr(sp2bp(sec, La, M))
            pos = (np.tile(np.arange(logm), La) + logm * np.repeat(amp_sections, logm)).astype(int)
            LLR[pos] = llr_a
            post[:(L - ls) * M] = sec[:(L - ls) * M]
        ber_amp.append(ber_from_LLRs(M, LLR, idx, total_bits))
        if record is not None:
            record["stages"].append(dict(app=app.copy(), it=it, amp_sections=list(amp_sections), LLR=LLR.copy()))
    return ber_amp, ber_ldpc, (L * logm - (nl - kl)) / n


# --------------------------------------------------------------------------- BPSK baseline
def sim_ldpc(lp, sigma, MIN_ERRORS=100, MAX_BLOCKS=400000, rng=np.random):
    """ldpc/sparc_ldpc.py:1064-1124 with awgn/ch2llr/bpsk :1049-1061."""
    if lp.r_ldpc not in ("1/2", "2/3", "3/4", "5/6", "0.45"):
        raise NameError("Rate unsupported")
    code = get_code(lp.standard, lp.r_ldpc, lp.z, lp.ptype)
    nbit = nblockerr = nblocks = 0
    ber = 0.0
    while nblockerr < MIN_ERRORS:
        if lp.standard in ("802.11n", "802.16"):
            x = code.encode(rng.randint(0, 2, code.K))
        else:
            x = np.zeros(code.N)
        yc = (1.0 - 2.0 * x) + sigma * rng.randn(len(x))
        app, _ = code.decode(2.0 / sigma ** 2 * yc, "sumprod2")
        e = int(np.sum(x != (app < 0.0)))
        nbit += e
        nblockerr += 1 if e else 0
        nblocks += 1
        ber = nbit / (nblocks * code.N)
        if nblocks >= MAX_BLOCKS:
            break
    return ber


# --------------------------------------------------------------------------- EXIT chart
def J_inverse(I):
    """ldpc/amp_exit.py:28-36."""
    assert 0 <= I <= 1
    if I == 1:
        I = 0.9999
    if I <= 0.3646:
        return 1.09542 * (I ** 2) + 0.214217 * I + 2.33727 * np.sqrt(I)
    return -0.706692 * np.log(0.386013 * (1 - I)) + 1.75017 * I


def gen_bits(length, rng=np.random):
    """ldpc/amp_exit.py:48-50."""
    return (rng.randint(0, 2, length) * -2) + 1


def prep_y(X, L, M, n, sigma_w, P, a=None, f=None, C=None, rng=np.random):
    """ldpc/amp_exit.py:125-160."""
    Pl = P / L * np.ones(L) if a is None else pa_parameterised(L, C, P, a, f)
    idx = bits2indices((X - 1) * -1 / 2, M)
    y, Ab, Az, ordering = _encode_and_channel(idx, L, M, n, Pl, sigma_w, rng)
    return y, Ab, Az, Pl, ordering


def calc_E(X, I_a, snr_dB, sp, threshold=0.5, rng=np.random, record=None):
    """ldpc/amp_exit.py:185-270 (without the CSV export)."""
    L, M, T = sp.L, sp.M, sp.t
    logm = int(np.log2(M))
    n = int(L * np.log2(M) / sp.r)
    sigma_w = np.sqrt(sp.p / 10 ** (snr_dB / 20))
    sigma_a = J_inverse(I_a)
    mu_a = sigma_a ** 2 / 2
    A = mu_a * X + rng.randn(len(X)) * sigma_a
    with np.errstate(over="ignore"):
        bw = 1 / (1 + np.exp(A))
    beta0 = bp2sp(bw, L, M)
    y, Ab, _, Pl, ordering = prep_y(X, L, M, n, sigma_w, sp.p, sp.a, sp.f, sp.C, rng)
    y_new, Ab_n, Az_n, amp_sections, La = hard_initialisation(beta0, L, M, n, ordering, y, Pl, Ab, threshold)
    E = A
    if La > 0:
        bT, _ = amp(y_new, Pl[amp_sections], La, M, T, Ab_n, Az_n)
        sec = bT.reshape(-1) / np.sqrt(n * np.repeat(Pl[amp_sections], M))
        Ea = bitwise_to_llr(sp2bp(sec, La, M))
        pos = (np.tile(np.arange(logm), La) + logm * np.repeat(amp_sections, logm)).astype(int)
        E[pos] = Ea
    np.clip(E, -55, 55, out=E)
    if record is not None:
        record.update(y=y.reshape(-1).copy(), amp_sections=list(amp_sections))
    return E


def hist_E(X, E, bin_number=500, max_bin=40, min_bin=-40):
    """ldpc/amp_exit.py:272-326 (no plotting)."""
    assert len(E) == len(X)
    bin_width = (max_bin - min_bin) / (bin_number - 1)
    edges = np.linspace(min_bin, max_bin, bin_number)
    PE_pos, _ = np.histogram(E[np.where(X == 1)[0]], bins=edges, density=True)
    PE_neg, _ = np.histogram(E[np.where(X == -1)[0]], bins=edges, density=True)
    mids = 0.5 * (edges[1:] + edges[:-1])
    mean_pos = np.average(mids, weights=PE_pos)
    mean_neg = np.average(mids, weights=PE_neg)
    var_pos = np.average((mids - mean_pos) ** 2, weights=PE_pos)
    var_neg = np.average((mids - mean_neg) ** 2, weights=PE_neg)
    return PE_pos, PE_neg, mean_pos, mean_neg, var_pos, var_neg, bin_width


def calc_I_e(PE_pos, PE_neg, bin_width):
    """ldpc/amp_exit.py:328-351 (+ remove_common_zeros :162-176)."""
    both = (PE_pos == 0) & (PE_neg == 0)
    PE_pos, PE_neg = PE_pos[~both], PE_neg[~both]
    with np.errstate(divide="ignore", invalid="ignore"):
        i_neg = PE_neg * np.log2(2 * PE_neg / (PE_neg + PE_pos))
        i_pos = PE_pos * np.log2(2 * PE_pos / (PE_neg + PE_pos))
    i_neg[np.isnan(i_neg)] = 0
    i_pos[np.isnan(i_pos)] = 0
    return 1 / 2 * (bin_width * np.sum(i_neg) + bin_width * np.sum(i_pos))
