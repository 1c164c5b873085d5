"""CPU-only checks: the C-ABI library loads and exports every symbol include/sparc_b200.h declares (no compute
calls), and the host-side logic (code tables, encoder, orderings, bit maps, power allocation) matches the
golden vectors of the unmodified reference."""
import ctypes
import hashlib
import os
import re

import numpy as np
import pytest

from conftest import ROOT, golden


def sha(a):
    return np.frombuffer(hashlib.sha256(np.ascontiguousarray(a).tobytes()).digest()[:8], dtype=np.uint64)[0]


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "sparc_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"^\s*(?:const\s+char\s*\*\s*|(?:int|long|double|void)\s+)(\w+)\s*\(", src, flags=re.M)))


def test_library_exports_every_declared_symbol():
    from sparc_ldpc_b200 import _lib
    assert os.path.isfile(_lib.LIB_PATH), "build with `make -C sparc_ldpc_b200/csrc`"
    lib = ctypes.CDLL(_lib.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 25 and {"sumprod", "sumprod2", "minsum", "Lxor", "Lxfb", "sb_amp_batch", "sb_bp_batch"} <= set(names)
    for n in names:
        assert hasattr(lib, n), "symbol %s declared in include/sparc_b200.h but not exported" % n
    # the ctypes binding covers the same set
    assert set(_lib.SIGNATURES) == set(names)
    assert _lib.lib().sb_version() >= 100


def test_product_never_imports_oracle():
    """The product path must not route through the CPU oracle."""
    pkg = os.path.join(ROOT, "sparc_ldpc_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, f)).read()
                assert "oracle" not in txt.replace("no CPU oracle", ""), f


def test_code_tables_match_reference():
    from sparc_ldpc_b200 import ldpc
    g = golden("ldpc")
    for i in range(int(g["n_codes"])):
        p = "k%d_" % i
        std, rate, z, pt = g[p + "name"]
        c = ldpc.code(str(std), str(rate), int(z), str(pt))
        assert sha(c.vdeg) == g[p + "vdeg_sha"] and sha(c.cdeg) == g[p + "cdeg_sha"]
        assert sha(c.intrlv) == g[p + "intrlv_sha"], (std, rate, z)
        if p + "intrlv" in g:
            assert np.array_equal(c.intrlv, g[p + "intrlv"])
        if p + "u" in g:
            assert np.array_equal(c.encode(g[p + "u"]), g[p + "x"])
    c = ldpc.code("802.16", "1/2", 81)  # ldpc/src/ldpc802.16.81.h:1-3
    assert (c.Nv, c.Nc, c.Nmsg) == (1944, 972, 6156)


@pytest.mark.parametrize("std,rate,z,pt", [("802.16", "1/2", 3, "A"), ("802.16", "2/3", 3, "A"), ("802.16", "2/3", 3, "B"),
                                           ("802.16", "3/4", 27, "A"), ("802.16", "3/4", 27, "B"), ("802.16", "5/6", 54, "A"),
                                           ("802.11n", "1/2", 27, "A"), ("802.11n", "2/3", 54, "A"), ("802.11n", "3/4", 81, "A"),
                                           ("802.11n", "5/6", 81, "A")])
def test_encoder_property(std, rate, z, pt):
    """ldpc/py/test_ldpc.py:44-58: 24 columns, degree sums, H x = 0 for random information words."""
    from sparc_ldpc_b200 import ldpc
    c = ldpc.code(std, rate, z, pt)
    assert len(c.proto[0]) == 24
    H = c.pcmat()
    assert np.sum(c.vdeg) == np.sum(c.cdeg) == np.sum(H) == len(c.intrlv)
    rs = np.random.RandomState(1)
    X = c.encode_batch(rs.randint(0, 2, (20, c.K)))
    assert np.count_nonzero(X.dot(H.T) % 2) == 0
    assert np.array_equal(X[3], c.encode(X[3][:c.K]))


def test_code_errors():
    from sparc_ldpc_b200 import ldpc
    for args in [("802.11n", "1/2", 28), ("802.16", "7/8", 4), ("802.16", "2/3", 4, "C"), ("nope", "1/2", 4)]:
        with pytest.raises(NameError):
            ldpc.code(*args)
    c = ldpc.code("802.16", "1/2", 3)
    with pytest.raises(NameError):
        c.encode(np.zeros(5, dtype=int))
    with pytest.raises(NameError):
        c.decode(np.zeros(5))


def test_ordering_and_host_maps(oracle):
    from sparc_ldpc_b200 import engine, sparc_ldpc as S
    g = golden("ops")
    assert sha(engine.make_ordering(256, 32, 1280)) == g["c4_ordering_sha"]
    assert np.array_equal(engine.make_ordering(8, 16, 24), g["a_ordering"])
    np.testing.assert_allclose(S.pa_parameterised(16, 1.2, 4.0, 0.7, 0.6), g["pa"], rtol=1e-14)
    with pytest.raises(IndexError):
        S.pa_parameterised(16, 1.2, 4.0, 1.0, 1.0)
    h = golden("handoff")
    for tag, M in (("m4", 4), ("m32", 32), ("m512", 512)):
        assert S.bits2indices(h[tag + "_bits"], M) == h[tag + "_idx"].tolist()


@pytest.mark.parametrize("shape", [(128, 4, 256), (64, 8, 192), (256, 32, 1280), (48, 512, 432), (33, 64, 200),
                                   (20, 512, 180), (40, 1024, 400)])
def test_fast_mode_tables_are_a_reordering_of_the_operator(shape):
    """FAST-mode lookup tables (scheduled for bank-conflict-free shared-memory gathers, csrc/sched.h) must list
    exactly the terms of the operator of sparc_ldpc.py:110-134; the builder is host code, so this runs without a
    GPU.  Also pins the bank-conflict model: the fold is (nearly) conflict-free."""
    import ctypes as ct
    from sparc_ldpc_b200 import _lib
    from sparc_ldpc_b200.engine import make_ordering
    L, M, n = shape
    o = make_ordering(L, M, n)
    st = (ct.c_long * 4)()
    rc = _lib.lib().sb_fast_tables_check(o.ctypes.data, L, M, n, st)
    assert rc == 0, _lib.lib().sb_last_error()
    assert st[0] > 0 and st[1] <= 1.25 * st[0]
    if st[2]:
        assert st[3] <= 2.0 * st[2]
    # a corrupted ordering entry (out of range) is rejected by the operator constructor's own checks, not here;
    # but a table built from a different ordering must not verify against this one -> exercised by the C check
    # itself (it compares every term with `ordering`)


@pytest.mark.parametrize("shape", [(48, 512, 432), (16, 512, 144), (64, 512, 4608), (24, 512, 2300), (40, 512, 1000)])
def test_pair_kernel_tables_are_a_reordering_of_the_operator(shape):
    """Tables of the two-codewords-per-CTA FAST kernel (csrc/amp2.cu): fold terms listed by sign half, gather terms
    with the sign in bit 15; must list exactly the operator's terms (host code, no GPU).  Shapes outside the
    kernel's range report 1 (no pair tables)."""
    import ctypes as ct
    from sparc_ldpc_b200 import _lib
    from sparc_ldpc_b200.engine import make_ordering
    L, M, n = shape
    o = make_ordering(L, M, n)
    st = (ct.c_long * 4)()
    rc = _lib.lib().sb_pair_tables_check(o.ctypes.data, L, M, n, st)
    assert rc == 0, _lib.lib().sb_last_error()
    assert st[0] > 0 and st[1] <= 1.3 * st[0] and st[3] <= 2.3 * st[2]
    # the fp64 tables of SB_AMP_F64: the same terms, scheduled for half-warp pools of 16 lanes x 16 eight-byte banks
    st2 = (ct.c_long * 4)()
    rc = _lib.lib().sb_pair_tables_check_f64(o.ctypes.data, L, M, n, st2)
    assert rc == 0, _lib.lib().sb_last_error()
    assert st2[0] == 2 * st[0] and st2[1] <= 1.3 * st2[0] and st2[3] <= 2.3 * st2[2]
    for L2, M2, n2 in ((20, 512, 180), (64, 256, 576), (16, 512, 6000)):   # L % 8, M != 512, w/M = 32
        o2 = make_ordering(L2, M2, n2)
        assert _lib.lib().sb_pair_tables_check(o2.ctypes.data, L2, M2, n2, None) == 1


def test_bench_helpers_without_a_gpu():
    """bench.py / tools used by it, host logic only: the union of overlapping launch intervals, the exponential power
    allocation of the C2 shape (waterfall(pa_param=True): a = f = r / C from the first grid point, sparc_ldpc.py:1201-1204),
    and the acceptance rule of the cliff-point comparison."""
    import importlib.util
    import sys
    from conftest import ROOT

    def load(name, path):
        spec = importlib.util.spec_from_file_location(name, path)
        mod = importlib.util.module_from_spec(spec)
        sys.modules[name] = mod
        spec.loader.exec_module(mod)
        return mod

    bench = load("bench_mod", os.path.join(ROOT, "bench.py"))

    class Ev:                                   # stands in for torch.cuda.Event: elapsed_time in ms
        def __init__(self, t):
            self.t = t

        def elapsed_time(self, other):
            return other.t - self.t

    e0 = Ev(0.0)
    pairs = [(Ev(1.0), Ev(4.0)), (Ev(3.0), Ev(6.0)), (Ev(10.0), Ev(11.0)), (Ev(10.5), Ev(10.75))]
    assert bench._union_ms(e0, pairs) == pytest.approx(6.0)
    assert bench._union_ms(e0, []) == 0.0
    assert bench.N == 4608 and bench.INFO_BITS == 3840 and abs(bench.SIGMA - 0.9963928922771221) < 1e-15

    shapes = load("bench_shapes_mod", os.path.join(ROOT, "tools", "bench_shapes.py"))
    pa = shapes.pa_exponential(512, 4.0, 1.0, 5.2)
    s0 = np.sqrt(4.0 / (10 ** (5.2 / 20) * 2 * 1.0))
    C = 0.5 * np.log2(1 + 4.0 / s0 ** 2)
    assert pa["C"] == pytest.approx(C) and pa["a"] == pytest.approx(1.0 / C) and pa["f"] == pa["a"] and pa["f"] < 1.0
    from sparc_ldpc_b200.sparc_ldpc import pa_parameterised
    Pl = pa_parameterised(512, pa["C"], 4.0, pa["a"], pa["f"])
    assert Pl.sum() == pytest.approx(4.0) and 3.0 < Pl[0] / Pl[-1] < 4.0 and np.all(np.diff(Pl) <= 1e-15)   # decaying, flat tail
    assert np.all(Pl[int(pa["f"] * 512):] == Pl[-1])

    cliff = load("cliff_point", os.path.join(ROOT, "tools", "cliff_point.py"))
    good = {"f64_vs_strict": {"converged_codewords": 3520, "converged_with_identical_decisions": 3520},
            "fast_vs_strict": {"converged_codewords": 3520, "converged_with_identical_decisions": 3519}}
    assert cliff.part_a_ok(good) and good["fast_vs_strict"]["near_tie_flips_among_converged"] == 1
    bad = {"f64_vs_strict": {"converged_codewords": 3520, "converged_with_identical_decisions": 3519},
           "fast_vs_strict": {"converged_codewords": 3520, "converged_with_identical_decisions": 3520}}
    assert not cliff.part_a_ok(bad)
    worse = {"f64_vs_strict": {"converged_codewords": 3520, "converged_with_identical_decisions": 3520},
             "fast_vs_strict": {"converged_codewords": 3520, "converged_with_identical_decisions": 3400}}
    assert not cliff.part_a_ok(worse)


def test_notebook_power_allocations():
    """sparc_amp.ipynb cells 9 / 13 (host functions of sparc_ldpc_b200.sparc_amp): sums, shapes, and the iterative
    allocation of the notebook's recorded run (L = 1024, sigma = 1, P = 15, R_PA = 1.4: blocks of one section)."""
    from sparc_ldpc_b200 import sparc_amp as SA
    pa = SA.pa_original(64, 1.5, 7.0)
    assert pa.sum() == pytest.approx(7.0) and np.all(np.diff(pa) < 0) and pa[0] / pa[-1] == pytest.approx(2 ** (2 * 1.5 * 63 / 64))
    PA = SA.pa_iterative(1024, 1024, 1.0, 15.0, 1.4)
    assert PA.sum() == pytest.approx(15.0) and np.all(np.diff(PA) <= 1e-18) and PA[-1] > 0
    # first section: P_block = 2 ln 2 (R_PA / L) (sigma^2 + P)
    assert PA[0] == pytest.approx(2 * np.log(2) * (1.4 / 1024) * 16.0)
    flat = np.flatnonzero(PA == PA[-1])
    assert 0 < flat[0] < 1024 and np.all(PA[flat[0]:] == PA[-1])          # the tail is spread evenly
    # with B < L the blocks are L // B sections wide
    PB = SA.pa_iterative(16, 4, 1.0, 15.0, 1.0)
    assert PB.sum() == pytest.approx(15.0) and PB[0] == PB[3]
