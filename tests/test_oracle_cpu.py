"""The CPU oracle against the golden vectors produced by the unmodified reference
(tests/golden/gen_golden.py) and the reference's own known answers.
Runs without a GPU.  The oracle makes the same numpy calls as the reference, so on the
machine that generated the goldens it is bit-identical; elsewhere numpy's SIMD exp/log
kernels may differ in the last ulp, hence the tight (not zero) tolerances on
transcendental-dependent quantities and exact comparison everywhere else."""
import hashlib
import os

import numpy as np
import pytest

from conftest import FLOW_CASES, ROOT, flat_result, golden

RT = dict(rtol=1e-9, atol=1e-12)


def sha(a):
    return np.frombuffer(hashlib.sha256(np.ascontiguousarray(a).tobytes()).digest()[:8], dtype=np.uint64)[0]


@pytest.mark.parametrize("tag", ["a", "b", "c1", "c"])
def test_operators_bit_exact(oracle, tag):
    g = golden("ops")
    L, M, n = (int(v) for v in g[tag + "_shape"])
    Ab, Az, ordering = oracle.sparc_transforms(L, M, n)
    assert np.array_equal(ordering, g[tag + "_ordering"])
    assert np.array_equal(Ab(g[tag + "_b"]).reshape(-1), g[tag + "_Ab"])
    assert np.array_equal(Az(g[tag + "_z"]).reshape(-1), g[tag + "_Az"])
    sub = g[tag + "_sub"]
    Ab2, Az2 = oracle.sparc_transforms_shorter(len(sub), M, n, ordering[sub, :])
    assert np.array_equal(Ab2(g[tag + "_b"][: len(sub) * M]).reshape(-1), g[tag + "_Ab_sub"])
    assert np.array_equal(Az2(g[tag + "_z"]).reshape(-1), g[tag + "_Az_sub"])


def test_ordering_fingerprints_and_pa(oracle):
    g = golden("ops")
    assert sha(oracle.make_ordering(256, 32, 1280)) == g["c4_ordering_sha"]
    assert sha(oracle.make_ordering(512, 512, 4608)) == g["c3_ordering_sha"]
    np.testing.assert_allclose(oracle.pa_parameterised(16, 1.2, 4.0, 0.7, 0.6), g["pa"], rtol=1e-14)
    with pytest.raises(IndexError):
        oracle.pa_parameterised(16, 1.2, 4.0, 1.0, 1.0)  # f >= 1: sparc_ldpc.py:184


@pytest.mark.parametrize("k", [0, 1, "w"])
def test_amp_trace_c1(oracle, k):
    g = golden("amp_small")
    L, M, P, T, n = 128, 4, 2.0, 64, 256
    Pl = P / L * np.ones(L)
    Ab, Az, _ = oracle.sparc_transforms(L, M, n)
    tr = []
    if k == "w":
        beta, t = oracle.amp(g["c1_0_y"].reshape(-1, 1), Pl, L, M, T, Ab, Az, g["c1_w_init"], trace=tr)
        p = "c1_w_"
    else:
        p = "c1_%d_" % k
        beta, t = oracle.amp(g[p + "y"].reshape(-1, 1), Pl, L, M, T, Ab, Az, trace=tr)
    assert t == int(g[p + "t"])
    np.testing.assert_allclose(beta.reshape(-1), g[p + "beta"], **RT)
    np.testing.assert_allclose(np.array([x[0] for x in tr]), g[p + "tau2"][: len(tr)], **RT)
    np.testing.assert_allclose(np.array([x[1] for x in tr]), g[p + "beta_trace"], **RT)


def test_amp_power_allocation(oracle):
    g = golden("amp_small")
    L, M, T = 32, 64, 64
    n = L * 6
    Ab, Az, _ = oracle.sparc_transforms(L, M, n)
    beta, t = oracle.amp(g["pa_y"].reshape(-1, 1), g["pa_Pl"], L, M, T, Ab, Az)
    assert t == int(g["pa_t"])
    np.testing.assert_allclose(beta.reshape(-1), g["pa_beta"], **RT)


def test_handoff_known_answers(oracle):
    g = golden("handoff")
    # ldpc/removed.py:40-49 and :203-204
    np.testing.assert_allclose(oracle.sp2bp(g["kat1_in"], 2, 4), [0.1, 0.15, 0.95, 0.9], rtol=1e-15)
    np.testing.assert_allclose(oracle.bp2sp(g["kat1_bp"], 2, 4),
                               [0.765, 0.135, 0.085, 0.015, 0.005, 0.045, 0.095, 0.855], rtol=1e-12)
    np.testing.assert_allclose(oracle.sp2bp(g["kat2_in"], 2, 4), [0.5, 0.7, 0, 0.4], rtol=1e-15)
    assert np.array_equal(oracle.bp2sp(g["kat1_bp"], 2, 4), g["kat1_back"])


@pytest.mark.parametrize("tag,L,M", [("m4", 16, 4), ("m32", 8, 32), ("m512", 3, 512)])
def test_handoff_maps(oracle, tag, L, M):
    g = golden("handoff")
    bp = oracle.sp2bp(g[tag + "_sec"], L, M)
    assert np.array_equal(bp, g[tag + "_bp"])
    assert np.array_equal(oracle.sp2bp_loops(g[tag + "_sec"], L, M), g[tag + "_bp"])
    llr = oracle.bitwise_to_llr(bp)
    np.testing.assert_allclose(llr, g[tag + "_llr"], rtol=1e-13)
    assert (llr == -np.finfo(float).max).sum() == (g[tag + "_llr"] == -np.finfo(float).max).sum() >= 1
    np.testing.assert_allclose(oracle.bp2sp(g[tag + "_bw"], L, M), g[tag + "_sp"], rtol=1e-14)
    assert oracle.bits2indices(g[tag + "_bits"], M) == g[tag + "_idx"].tolist()
    assert oracle.ber_from_LLRs(M, g[tag + "_llr"], g[tag + "_idx"].tolist(), L * int(np.log2(M))) == float(g[tag + "_ber"])


def test_ldpc_tables_encode_decode(oracle):
    g = golden("ldpc")
    for i in range(int(g["n_codes"])):
        p = "k%d_" % i
        std, rate, z, pt = g[p + "name"]
        c = oracle.Code(str(std), str(rate), int(z), str(pt))
        assert sha(c.vdeg) == g[p + "vdeg_sha"] and sha(c.cdeg) == g[p + "cdeg_sha"]
        assert sha(c.intrlv) == g[p + "intrlv_sha"], (std, rate, z)
        if p + "u" in g:
            x = c.encode(g[p + "u"])
            assert np.array_equal(x, g[p + "x"])
            assert not np.any(c.pcmat().dot(x) % 2) if c.N <= 2000 else True
        for j in range(3):
            app, it = c.decode(g[p + "ch"][j])
            assert it == g[p + "it"][j]
            np.testing.assert_allclose(app, g[p + "app"][j], rtol=1e-9, atol=1e-9)
        if p + "app_sumprod" in g:
            app, it = c.decode(g[p + "ch"][1], "sumprod")
            assert it == g[p + "it_sumprod"]
            np.testing.assert_allclose(app, g[p + "app_sumprod"], rtol=1e-9, atol=1e-9)
        if p + "app_minsum" in g:
            app, it = c.decode(g[p + "ch"][1], "minsum", 0.7)
            assert it == g[p + "it_minsum"]
            np.testing.assert_allclose(app, g[p + "app_minsum"], rtol=1e-9, atol=1e-9)
    out = np.array([oracle.Lxor(a, b) for a, b in g["lxor_in"]])
    np.testing.assert_allclose(out, g["lxor_out"], rtol=1e-14, equal_nan=True)
    tot, ext = oracle.Lxfb(g["lxfb_in"])
    np.testing.assert_allclose(tot, g["lxfb_tot"], rtol=1e-14)
    np.testing.assert_allclose(ext, g["lxfb_ext"], rtol=1e-14)


def test_reference_header_fixture(oracle):
    """ldpc/src/ldpc802.16.81.h: Nv=1944, Nc=972, Nmsg=6156 == prepare_decoder("802.16","1/2",81)."""
    c = oracle.Code("802.16", "1/2", 81)
    assert (c.Nv, c.Nc, c.Nmsg) == (1944, 972, 6156)


@pytest.mark.parametrize("std,rate,z,pt", [("802.16", "1/2", 3, "A"), ("802.16", "2/3", 3, "B"), ("802.16", "3/4", 27, "A"),
                                           ("802.16", "5/6", 27, "A"), ("802.11n", "2/3", 27, "A"), ("802.11n", "3/4", 54, "A")])
def test_reference_pytest_property(oracle, std, rate, z, pt):
    """ldpc/py/test_ldpc.py:44-65: H x = 0 and a noiseless word decodes with it == 0."""
    c = oracle.Code(std, rate, z, pt)
    assert c.proto.shape[1] == 24
    H = c.pcmat()
    assert c.vdeg.sum() == c.cdeg.sum() == H.sum() == len(c.intrlv)
    rs = np.random.RandomState(0)
    for _ in range(5):
        u = rs.randint(0, 2, c.K)
        x = c.encode(u)
        assert np.count_nonzero(H.dot(x) % 2) == 0
        app, it = c.decode(np.array(10 * (.5 - x), dtype=float))
        assert it == 0 and np.array_equal((app < 0).astype(int), x)


def test_bp_restatement_equals_reference_library(oracle):
    """oracle.c's sumprod2 against the reference's own c_ldpc.c compiled into oracle/_ref (when present)."""
    import ctypes

    path = os.path.join(ROOT, "oracle", "_ref", "bin", "c_ldpc.so")
    if not os.path.isfile(path):
        pytest.skip("oracle/_ref/bin/c_ldpc.so not built")
    ref = ctypes.CDLL(path)
    c = oracle.Code("802.16", "5/6", 24)
    rs = np.random.RandomState(5)
    for s in (0.5, 0.6, 0.75):
        x = c.encode(rs.randint(0, 2, c.K))
        ch = 2 / s ** 2 * (1 - 2.0 * x + s * rs.randn(c.N))
        app_r = np.zeros(c.N)
        it_r = ref.sumprod2(ch.ctypes.data_as(ctypes.c_void_p), c.vdeg.ctypes.data_as(ctypes.c_void_p),
                            c.cdeg.ctypes.data_as(ctypes.c_void_p), c.intrlv.ctypes.data_as(ctypes.c_void_p),
                            c.Nv, c.Nc, c.Nmsg, app_r.ctypes.data_as(ctypes.c_void_p))
        app, it = c.decode(ch)
        assert it == it_r and np.array_equal(app, app_r)


@pytest.mark.parametrize("case", FLOW_CASES, ids=[c[0] for c in FLOW_CASES])
def test_link_sims(oracle, case):
    tag, fn, spk, lpk, kw, reps = case
    g = golden("flows_small")
    rng = np.random.RandomState(int(g[tag + "_seed"]))
    sp = oracle.SPARCParams(**spk)
    lp = None if lpk is None else oracle.LDPCParams(*lpk)
    rows = []
    for _ in range(reps):
        if fn == "amp_ldpc_sim":
            res = oracle.amp_ldpc_sim(sp, lp, rng=rng)
        elif fn == "soft_amp_ldpc_sim":
            res = oracle.soft_amp_ldpc_sim(sp, lp, kw["soft_iter"], rng=rng)
        elif fn == "hardinitbeta_amp_ldpc_sim":
            res = oracle.hardinitbeta_amp_ldpc_sim(sp, lp, rng=rng)
        else:
            res = oracle.soft_amp_ldpc_hardinit(sp, lp, kw["soft_iter"], kw["threshold"], rng=rng)
        rows.append(flat_result(res))
    np.testing.assert_array_equal(np.array(rows), g[tag + "_res"])


def test_exit_chain(oracle):
    g = golden("exit")
    rng = np.random.RandomState(77)
    sp = oracle.SPARCParams(L=64, M=8, sigma=None, p=4, r=1, t=64)
    for k, (I_a, snr, thr) in enumerate(g["meta"]):
        X = oracle.gen_bits(64 * 3, rng)
        assert np.array_equal(X, g["X"][k])
        E = oracle.calc_E(X, I_a, snr, sp, threshold=thr, rng=rng)
        np.testing.assert_allclose(E, g["E"][k], rtol=1e-9, atol=1e-9)
        h = oracle.hist_E(X, E, 60, 60, -60)
        np.testing.assert_allclose(oracle.calc_I_e(h[0], h[1], h[6]), g["I_e"][k], rtol=1e-9)
        if I_a == 0.5:
            np.testing.assert_allclose(h[0], g["pe_pos"], rtol=1e-12)
            np.testing.assert_allclose(h[1], g["pe_neg"], rtol=1e-12)
    np.testing.assert_allclose([oracle.J_inverse(v) for v in (0.0, 0.2, 0.3646, 0.5, 0.99, 1.0)], g["J_inv"], rtol=1e-14)
