"""Generate the committed golden vectors by running the UNMODIFIED reference.

Run in the build container only (needs /root/reference and `make -C oracle`):

    python tests/golden/gen_golden.py            # writes tests/golden/*.npz

Nothing is copied from the reference: its modules are imported through
oracle/ref_harness.py (shims + monkeypatches documented there) and driven with
seeded inputs.  Per-iteration AMP state is captured without touching the
reference's `amp` by wrapping the Ab/Az closures it is handed: Az is called with
z_t (so tau_t^2 is recomputed exactly as sparc_ldpc.py:203 does) and Ab with
beta_{t+1}.  AMP / BP calls made inside the link simulations are captured by
wrapping the module-level `amp` and `ldpc.code.decode` names.
"""
import hashlib
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import ref_harness as rh  # noqa: E402

warnings.simplefilter("ignore")
sl, ae, at, ldpc = rh.load_reference()
OUT = HERE


def save(name, **arrs):
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **arrs)
    print("%-24s %8.1f kB" % (name, os.path.getsize(path) / 1e3))


def sha(a):
    return np.frombuffer(hashlib.sha256(np.ascontiguousarray(a).tobytes()).digest()[:8], dtype=np.uint64)[0]


def traced_amp(y, Pl, L, M, T, Ab, Az, beta0=None):
    zs, bs = [], []

    def Ab_t(b):
        bs.append(np.array(b, dtype=np.float64).reshape(-1).copy())
        return Ab(b)

    def Az_t(z):
        zs.append(np.array(z, dtype=np.float64).reshape(-1).copy())
        return Az(z)

    beta, t = at.amp_test(y, None, Pl, L, M, T, Ab_t, Az_t, beta0)
    bs = bs[1:]  # first Ab call is z = y - Ab(beta0) (the harness always passes an explicit beta0)
    n = y.size
    tau2 = np.array([np.sqrt(np.sum(z.reshape(-1, 1) ** 2) / n) ** 2 for z in zs])
    return beta.reshape(-1), t, tau2, zs, bs


# ---------------------------------------------------------------- operators
def gen_ops():
    d = {}
    for tag, (L, M, n) in {"a": (8, 16, 24), "b": (6, 8, 40), "c1": (128, 4, 256), "c": (5, 64, 30)}.items():
        Ab, Az, ordering = sl.sparc_transforms(L, M, n)
        rs = np.random.RandomState(7)
        b = rs.randn(L * M)
        z = rs.randn(n)
        d[tag + "_shape"] = np.array([L, M, n])
        d[tag + "_ordering"] = ordering
        d[tag + "_b"], d[tag + "_z"] = b, z
        d[tag + "_Ab"], d[tag + "_Az"] = Ab(b).reshape(-1), Az(z).reshape(-1)
        sub = np.array([1, 3, 4]) if L > 4 else np.arange(L)
        Ab2, Az2 = sl.sparc_transforms_shorter(len(sub), M, n, ordering[sub, :])
        d[tag + "_sub"] = sub
        d[tag + "_Ab_sub"] = Ab2(b[: len(sub) * M]).reshape(-1)
        d[tag + "_Az_sub"] = Az2(z).reshape(-1)
    # full-size ordering fingerprints (table itself is 9.4 MB)
    for tag, (L, M, n) in {"c3": (512, 512, 4608), "c4": (256, 32, 1280)}.items():
        _, _, ordering = sl.sparc_transforms(L, M, n)
        d[tag + "_ordering_sha"] = sha(ordering)
        d[tag + "_ordering_head"] = ordering[:2, :16]
    d["pa"] = sl.pa_parameterised(16, 1.2, 4.0, 0.7, 0.6)
    save("ops", **d)


# ---------------------------------------------------------------- AMP traces
def gen_amp():
    d = {}
    # C1: L=128 M=4 r=1 P=2, two noise levels, plus a warm-started run
    L, M, P, r, T = 128, 4, 2.0, 1, 64
    n = int(L * np.log2(M) / r)
    Pl = P / L * np.ones(L)
    Ab, Az, _ = sl.sparc_transforms(L, M, n)
    for k, sigma in enumerate([0.708, 0.45]):
        np.random.seed(100 + k)
        bits = np.random.randint(0, 2, int(L * np.log2(M))).tolist()
        idx = sl.bits2indices(bits, M)
        b0 = np.zeros((L * M, 1))
        for l in range(L):
            b0[l * M + idx[l]] = np.sqrt(n * Pl[l])
        y = (Ab(b0) + np.random.randn(n, 1) * sigma).reshape(-1, 1)
        beta, t, tau2, zs, bs = traced_amp(y, Pl, L, M, T, Ab, Az)
        p = "c1_%d_" % k
        d[p + "idx"], d[p + "y"], d[p + "beta"], d[p + "t"] = np.array(idx), y.reshape(-1), beta, t
        d[p + "tau2"], d[p + "z_trace"], d[p + "beta_trace"] = tau2, np.array(zs), np.array(bs)
        if k == 0:
            init = (0.5 * b0 + 0.5 * np.sqrt(n * P / L) / M).reshape(-1)
            beta, t, tau2, zs, bs = traced_amp(y, Pl, L, M, T, Ab, Az, init)
            d["c1_w_init"], d["c1_w_beta"], d["c1_w_t"], d["c1_w_tau2"] = init, beta, t, tau2
            d["c1_w_beta_trace"] = np.array(bs)
    # mid size with exponential power allocation: L=32 M=64
    L, M, P, r, T = 32, 64, 4.0, 1, 64
    n = int(L * np.log2(M) / r)
    sigma = 0.7
    C = 0.5 * np.log2(1 + P / sigma ** 2)
    Pl = sl.pa_parameterised(L, C, P, r / C, r / C)
    Ab, Az, _ = sl.sparc_transforms(L, M, n)
    np.random.seed(5)
    idx = sl.bits2indices(np.random.randint(0, 2, int(L * np.log2(M))).tolist(), M)
    b0 = np.zeros((L * M, 1))
    for l in range(L):
        b0[l * M + idx[l]] = np.sqrt(n * Pl[l])
    y = (Ab(b0) + np.random.randn(n, 1) * sigma).reshape(-1, 1)
    beta, t, tau2, zs, bs = traced_amp(y, Pl, L, M, T, Ab, Az)
    d["pa_Pl"], d["pa_idx"], d["pa_y"], d["pa_beta"], d["pa_t"], d["pa_tau2"] = Pl, np.array(idx), y.reshape(-1), beta, t, tau2
    d["pa_beta_trace"] = np.array(bs)
    save("amp_small", **d)

    # C3 shape: L=M=512, r=1, P=4, one plain AMP decode at ref-dB 8 (sigma from R=1)
    d = {}
    L, M, P, r, T = 512, 512, 4.0, 1, 64
    n = int(L * np.log2(M) / r)
    Pl = P / L * np.ones(L)
    Ab, Az, _ = sl.sparc_transforms(L, M, n)
    for k, ebno_db in enumerate([8.0, 6.5]):
        sigma = np.sqrt(P / (10 ** (ebno_db / 20) * 2 * 1.0))
        np.random.seed(200 + k)
        idx = sl.bits2indices(np.random.randint(0, 2, int(L * np.log2(M))).tolist(), M)
        b0 = np.zeros((L * M, 1))
        for l in range(L):
            b0[l * M + idx[l]] = np.sqrt(n * Pl[l])
        y = (Ab(b0) + np.random.randn(n, 1) * sigma).reshape(-1, 1)
        beta, t, tau2, zs, bs = traced_amp(y, Pl, L, M, T, Ab, Az)
        p = "c3_%d_" % k
        B = np.array(bs).reshape(len(bs), L, M)
        d[p + "sigma"], d[p + "idx"], d[p + "y"], d[p + "t"], d[p + "tau2"] = sigma, np.array(idx), y.reshape(-1), t, tau2
        d[p + "argmax_trace"] = B.argmax(axis=2).astype(np.int16)
        d[p + "max_trace"] = B.max(axis=2)
        d[p + "sumsq_trace"] = np.array([np.sum(b ** 2) for b in bs])
        d[p + "z_last"] = zs[-1]
        post = beta / np.sqrt(n * np.repeat(Pl, M))
        d[p + "bitwise"] = sl.sp2bp(post, L, M)
        d[p + "beta_sha"] = sha(beta)
    save("amp_c3", **d)


# ---------------------------------------------------------------- handoff maps
def gen_handoff():
    d = {}
    # the reference's own inline known answers (ldpc/removed.py:40-49, :203-204)
    k1 = np.array([0.8, 0.1, 0.05, 0.05, 0, 0.05, 0.1, 0.85])
    d["kat1_in"], d["kat1_bp"] = k1, sl.sp2bp(k1, 2, 4)
    d["kat1_back"] = sl.bp2sp(d["kat1_bp"], 2, 4)
    k2 = np.array([0.2, 0.3, 0.1, 0.4, 0.6, 0.4, 0, 0])
    d["kat2_in"], d["kat2_bp"] = k2, sl.sp2bp(k2, 2, 4)
    rs = np.random.RandomState(11)
    for tag, (L, M) in {"m4": (16, 4), "m32": (8, 32), "m512": (3, 512)}.items():
        logm = int(np.log2(M))
        sec = rs.dirichlet(np.ones(M) * 0.05, size=L).reshape(-1)
        sec[:M] = 0
        sec[M // 2 + 1] = 1.0  # a saturated section on an odd index: p == 1 exactly
        bp = sl.sp2bp(sec, L, M)
        with np.errstate(all="ignore"):
            llr = np.nan_to_num(np.log(1 - bp) - np.log(bp))
        app = rs.randn(L * logm) * 6
        app[:3] = [800.0, -800.0, 0.0]
        with np.errstate(all="ignore"):
            bw = 1 / (1 + np.exp(app))
        d[tag + "_sec"], d[tag + "_bp"], d[tag + "_llr"] = sec, bp, llr
        d[tag + "_app"], d[tag + "_bw"], d[tag + "_sp"] = app, bw, sl.bp2sp(bw, L, M)
        bits = rs.randint(0, 2, L * logm)
        d[tag + "_bits"], d[tag + "_idx"] = bits, np.array(sl.bits2indices(bits.tolist(), M))
        d[tag + "_ber"] = sl.ber_from_LLRs(M, llr, d[tag + "_idx"].tolist(), L * logm)
    save("handoff", **d)


# ---------------------------------------------------------------- LDPC
def gen_ldpc():
    d = {}
    codes = [("802.16", "1/2", 3, "A"), ("802.16", "5/6", 8, "A"), ("802.16", "2/3", 5, "B"), ("802.16", "3/4", 4, "A"),
             ("802.11n", "1/2", 27, "A"), ("802.11n", "5/6", 27, "A"), ("2_7_12_good", "1/2", 4, "A"),
             ("2_5_12_good_threshold08", "3/8", 3, "A"), ("802.16", "5/6", 192, "A"), ("802.16", "1/2", 81, "A")]
    d["n_codes"] = len(codes)
    rs = np.random.RandomState(3)
    for i, (std, rate, z, pt) in enumerate(codes):
        c = ldpc.code(std, rate, z, pt)
        p = "k%d_" % i
        d[p + "name"] = np.array([std, rate, str(z), pt])
        big = c.Nmsg > 3000
        d[p + "vdeg_sha"], d[p + "cdeg_sha"], d[p + "intrlv_sha"] = sha(c.vdeg), sha(c.cdeg), sha(c.intrlv)
        if not big:
            d[p + "vdeg"], d[p + "cdeg"], d[p + "intrlv"] = c.vdeg, c.cdeg, c.intrlv
        if std in ("802.16", "802.11n"):
            u = rs.randint(0, 2, c.K)
            x = c.encode(u)
            d[p + "u"], d[p + "x"] = u.astype(np.int8), x.astype(np.int8)
        else:
            x = np.zeros(c.N, dtype=int)
        # decode cases: noiseless (it == 0), moderate noise (converges), heavy noise (hits 200)
        sig = {"1/2": (0.75, 1.1), "2/3": (0.6, 0.9), "3/4": (0.55, 0.8), "5/6": (0.5, 0.7), "3/8": (0.8, 1.3)}[rate]
        chs, apps, its = [], [], []
        for s in (None,) + sig:
            if s is None:
                ch = np.array(10 * (.5 - x), dtype=float)
            else:
                ch = 2.0 / s ** 2 * ((1.0 - 2.0 * x) + s * rs.randn(c.N))
            app, it = c.decode(ch, "sumprod2")
            chs.append(ch); apps.append(app.copy()); its.append(it)
        d[p + "ch"], d[p + "app"], d[p + "it"] = np.array(chs), np.array(apps), np.array(its)
        if i in (1, 4):
            app, it = c.decode(chs[1], "sumprod")
            d[p + "app_sumprod"], d[p + "it_sumprod"] = app.copy(), it
        if i == 1:  # check-regular code: the reference's minsum indexing bug is dormant
            app, it = c.decode(chs[1], "minsum", 0.7)
            d[p + "app_minsum"], d[p + "it_minsum"] = app.copy(), it
    c = ldpc.code("802.16", "1/2", 3)
    d["lxor_in"] = np.array([[1.5, -2.25], [0.0, 3.0], [-0.0, 3.0], [40.0, 41.0], [-1e308, 1e308], [1.797e308, 1.797e308]])
    d["lxor_out"] = np.array([c.Lxor(a, b) for a, b in d["lxor_in"]])
    d["lxor_out_nocorr"] = np.array([c.Lxor(a, b, 0) for a, b in d["lxor_in"]])
    Lin = rs.randn(7) * 3
    tot, ext = c.Lxfb(Lin)
    d["lxfb_in"], d["lxfb_tot"], d["lxfb_ext"] = Lin, tot, ext
    save("ldpc", **d)


# ---------------------------------------------------------------- link simulations
class Capture:
    """Records every amp() and code.decode() call the reference makes inside a link sim."""

    def __init__(self):
        self.amp_calls, self.dec_calls = [], []
        self._amp, self._dec = sl.amp, ldpc.code.decode
        cap = self

        def amp(y, s_n, Pl, L, M, T, Ab, Az, beta=None):
            out = cap._amp(y, s_n, Pl, L, M, T, Ab, Az, beta)
            cap.amp_calls.append(dict(y=np.array(y).reshape(-1).copy(), Pl=np.array(Pl).copy(), L=L,
                                      init=None if beta is None else np.array(beta).reshape(-1).copy(),
                                      out=np.array(out).reshape(-1).copy()))
            return out

        def decode(self_, ch, *a, **k):
            app, it = cap._dec(self_, ch, *a, **k)
            cap.dec_calls.append(dict(ch=np.array(ch).copy(), app=app.copy(), it=it))
            return app, it

        self.amp, self.decode = amp, decode

    def __enter__(self):
        sl.amp = ae.amp = self.amp
        ldpc.code.decode = self.decode
        return self

    def __exit__(self, *a):
        sl.amp = ae.amp = self._amp
        ldpc.code.decode = self._dec


def _flat(x):
    return np.array([np.nan if v is None else v for v in np.atleast_1d(np.array(x, dtype=object)).tolist()], dtype=float)


def gen_flows():
    d = {}
    cases = [
        ("plain_c1", "amp_ldpc_sim", dict(L=128, M=4, sigma=0.708, p=2, r=1, t=64), None, {}, 4),
        ("orig_s", "amp_ldpc_sim", dict(L=64, M=8, sigma=0.8, p=4, r=1, t=64), ("802.16", "5/6", 4), {}, 4),
        ("soft_s", "soft_amp_ldpc_sim", dict(L=64, M=8, sigma=0.8, p=4, r=1, t=64), ("802.16", "5/6", 8), dict(soft_iter=2), 4),
        ("hard_s", "hardinitbeta_amp_ldpc_sim", dict(L=64, M=8, sigma=0.8, p=4, r=1, t=64), ("802.16", "5/6", 8), {}, 4),
        ("thr_s", "soft_amp_ldpc_hardinit", dict(L=64, M=8, sigma=0.85, p=4, r=1, t=64), ("802.16", "5/6", 8), dict(soft_iter=3, threshold=0.6), 4),
        ("thr_m32", "soft_amp_ldpc_hardinit", dict(L=96, M=32, sigma=1.0, p=4, r=1, t=64), ("802.16", "1/2", 20), dict(soft_iter=3, threshold=0.7), 3),
        ("soft_m32", "soft_amp_ldpc_sim", dict(L=96, M=32, sigma=1.0, p=4, r=1, t=64), ("802.16", "1/2", 20), dict(soft_iter=2), 3),
    ]
    for tag, fn, spk, lpk, kw, reps in cases:
        np.random.seed(sum(map(ord, tag)))
        sp = sl.SPARCParams(**spk)
        lp = None if lpk is None else sl.LDPCParams(*lpk)
        rows = []
        with Capture() as cap:
            for _ in range(reps):
                if lp is None:
                    res = sl.amp_ldpc_sim(sp)
                elif fn == "amp_ldpc_sim":
                    res = sl.amp_ldpc_sim(sp, lp)
                elif fn == "soft_amp_ldpc_sim":
                    res = sl.soft_amp_ldpc_sim(sp, lp, kw["soft_iter"])
                elif fn == "hardinitbeta_amp_ldpc_sim":
                    res = sl.hardinitbeta_amp_ldpc_sim(sp, lp)
                else:
                    res = sl.soft_amp_ldpc_hardinit(sp, lp, kw["soft_iter"], kw["threshold"])
                rows.append(np.concatenate([_flat(r) for r in res]))
        d[tag + "_seed"] = sum(map(ord, tag))
        d[tag + "_res"] = np.array(rows)
        d[tag + "_n_amp"] = len(cap.amp_calls)
        d[tag + "_n_dec"] = len(cap.dec_calls)
        # BP iteration counts of every decode call, per repetition (200 = did not converge: chaotic orbit)
        d[tag + "_its"] = np.array([c["it"] for c in cap.dec_calls]).reshape(reps, -1) if cap.dec_calls else np.zeros((reps, 0), dtype=int)
        # first codeword's captured calls (small sizes: keep everything)
        for j, c in enumerate(cap.amp_calls[: len(cap.amp_calls) // reps]):
            d["%s_amp%d_y" % (tag, j)], d["%s_amp%d_out" % (tag, j)] = c["y"], c["out"]
            d["%s_amp%d_L" % (tag, j)] = c["L"]
            if c["init"] is not None:
                d["%s_amp%d_init" % (tag, j)] = c["init"]
        for j, c in enumerate(cap.dec_calls[: len(cap.dec_calls) // reps]):
            d["%s_dec%d_ch" % (tag, j)], d["%s_dec%d_app" % (tag, j)], d["%s_dec%d_it" % (tag, j)] = c["ch"], c["app"], c["it"]
    save("flows_small", **d)

    # one C3 codeword per flow at full size (L=M=512, 802.16 5/6 z=192)
    d = {}
    lp = sl.LDPCParams("802.16", "5/6", 192)
    for tag, fn, sigma, kw in [("soft", "soft", 1.02, dict(soft_iter=2)), ("hard", "hard", 1.02, {}),
                               ("thr", "thr", 1.05, dict(soft_iter=2, threshold=0.6))]:
        np.random.seed(sum(map(ord, tag)) + 1000)
        sp = sl.SPARCParams(L=512, M=512, sigma=sigma, p=4, r=1, t=64)
        with Capture() as cap:
            if fn == "soft":
                res = sl.soft_amp_ldpc_sim(sp, lp, 2)
            elif fn == "hard":
                res = sl.hardinitbeta_amp_ldpc_sim(sp, lp)
            else:
                res = sl.soft_amp_ldpc_hardinit(sp, lp, 2, 0.6)
        d[tag + "_seed"], d[tag + "_sigma"] = sum(map(ord, tag)) + 1000, sigma
        d[tag + "_res"] = np.concatenate([_flat(r) for r in res])
        d[tag + "_y"] = cap.amp_calls[0]["y"]
        n = 4608
        for j, c in enumerate(cap.amp_calls):
            L = c["L"]
            out = c["out"].reshape(L, 512)
            d["%s_amp%d_L" % (tag, j)] = L
            d["%s_amp%d_argmax" % (tag, j)] = out.argmax(axis=1).astype(np.int16)
            d["%s_amp%d_max" % (tag, j)] = out.max(axis=1)
            d["%s_amp%d_bitwise" % (tag, j)] = sl.sp2bp(c["out"] / np.sqrt(n * 4 / 512), L, 512)
        for j, c in enumerate(cap.dec_calls):
            d["%s_dec%d_ch" % (tag, j)], d["%s_dec%d_app" % (tag, j)], d["%s_dec%d_it" % (tag, j)] = c["ch"], c["app"], c["it"]
    save("flows_c3", **d)


# ---------------------------------------------------------------- EXIT chart
def gen_exit():
    d = {}
    sp = sl.SPARCParams(L=64, M=8, sigma=None, p=4, r=1, t=64)
    np.random.seed(77)
    rows_E, rows_X, rows_Ie, meta = [], [], [], []
    for I_a, snr_dB, thr in [(0.0, 10.0, 0.7), (0.5, 10.0, 0.7), (0.9, 12.0, 0.85), (0.99, 13.0, 0.95)]:
        X = ae.gen_bits(64 * 3)
        E = ae.calc_E(X, I_a, snr_dB, sp, threshold=thr)
        h = ae.hist_E(X, E, bin_number=60, max_bin=60, min_bin=-60)
        rows_X.append(X.copy()); rows_E.append(E.copy()); meta.append([I_a, snr_dB, thr])
        rows_Ie.append(ae.calc_I_e(h[0], h[1], h[6]))
        if I_a == 0.5:
            d["pe_pos"], d["pe_neg"], d["stats"] = h[0], h[1], np.array(h[2:])
    d["X"], d["E"], d["I_e"], d["meta"] = np.array(rows_X), np.array(rows_E), np.array(rows_Ie), np.array(meta)
    d["J_inv"] = np.array([ae.J_inverse(v) for v in (0.0, 0.2, 0.3646, 0.5, 0.99, 1.0)])
    # C4 shape sample: L=256 M=32
    sp = sl.SPARCParams(L=256, M=32, sigma=None, p=4, r=1, t=64)
    np.random.seed(78)
    X = ae.gen_bits(256 * 5)
    E = ae.calc_E(X, 0.66, 11.0, sp, threshold=0.7)
    h = ae.hist_E(X, E, bin_number=350, max_bin=60, min_bin=-60)
    d["c4_X"], d["c4_E"], d["c4_I_e"], d["c4_bw"] = X, E, ae.calc_I_e(h[0], h[1], h[6]), h[6]
    save("exit", **d)


if __name__ == "__main__":
    which = sys.argv[1:] or ["ops", "amp", "handoff", "ldpc", "flows", "exit"]
    for w in which:
        globals()["gen_" + w]()
