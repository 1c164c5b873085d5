"""Run the UNMODIFIED reference (/root/reference/ldpc) in this container.

TEST INFRASTRUCTURE ONLY -- used by tests/golden/gen_golden.py (golden-vector
generation) and by tests that cross-check the oracle restatement against the
live reference when the reference tree is mounted.  Nothing here may be
imported from sparc_ldpc_b200/.  `/root/reference` does not exist on the GPU
box, so GPU tests, smoke() and bench.py never import this module.

What is needed to import the reference without editing it:
  * shim modules for pylab / matplotlib / bitarray (oracle/shims),
  * `import amp_exit` BEFORE `sparc_ldpc` (circular import, sparc_ldpc.py:9 <->
    amp_exit.py:11),
  * cwd containing bin/c_ldpc.so (ldpc/py/ldpc.py:859 dlopens './bin/c_ldpc.so'),
    here oracle/_ref/ built by oracle/Makefile from the reference's c_ldpc.c,
  * three monkeypatches for numpy >= 2: `fht_inplace` -> compiled transcription
    of sparc_ldpc.py:19-29 (bit-identical, ~300x faster than the Python loop),
    `amp` wrapper that turns the `[None]` sentinel (sparc_ldpc.py:192, broken
    under numpy 2) into an explicit zero beta (bit-identical: y - Ab(0) == y),
    and a tolerant `np.set_printoptions` (amp_exit.py:261 passes threshold=nan).
"""
import ctypes
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = os.environ.get("SPARC_REF_ROOT", "/root/reference")
REF_LDPC = os.path.join(REF_ROOT, "ldpc")
REF_CWD = os.path.join(HERE, "_ref")

_loaded = None


def reference_available():
    return os.path.isfile(os.path.join(REF_LDPC, "sparc_ldpc.py")) and os.path.isfile(
        os.path.join(REF_CWD, "bin", "c_ldpc.so")
    )


def _c_fht():
    lib = ctypes.CDLL(os.path.join(HERE, "_build", "liboracle.so"))
    lib.orc_fht_inplace.argtypes = [ctypes.c_void_p, ctypes.c_long]
    lib.orc_fht_inplace.restype = None

    def fht_inplace(x):
        assert x.dtype == np.float64 and x.flags.c_contiguous
        lib.orc_fht_inplace(x.ctypes.data, x.size)

    return fht_inplace


def load_reference(fast_fht=True):
    """Returns (sparc_ldpc, amp_exit, amp_test, ldpc) reference modules. Changes cwd."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not reference_available():
        raise RuntimeError("reference tree or oracle/_ref/bin/c_ldpc.so missing (run `make -C oracle`)")
    os.chdir(REF_CWD)
    for p in (REF_LDPC, os.path.join(HERE, "shims")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import warnings

    _orig_spo = np.set_printoptions

    def _tolerant_spo(*a, **k):
        thr = k.get("threshold", None)
        if isinstance(thr, float) and thr != thr:
            k.pop("threshold")
        return _orig_spo(*a, **k)

    np.set_printoptions = _tolerant_spo
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        import amp_exit as ae  # noqa: E402  (must come first)
        import sparc_ldpc as sl  # noqa: E402
        import amp_test as at  # noqa: E402
        import py.ldpc as ldpc  # noqa: E402

    if fast_fht:
        sl.fht_inplace = _c_fht()

    def _wrap(orig):
        def amp(y, s_n, Pl, L, M, T, Ab, Az, beta=None):
            if beta is None or (isinstance(beta, np.ndarray) and beta.dtype == object):
                beta = np.zeros((L * M, 1))
            return orig(y, s_n, Pl, L, M, T, Ab, Az, beta)

        return amp

    sl_amp = _wrap(sl.amp)
    sl.amp = sl_amp
    ae.amp = sl_amp
    at.amp = sl_amp
    at.amp_test = _wrap(at.amp_test)
    _loaded = (sl, ae, at, ldpc)
    return _loaded
