import os
import sys
import warnings

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    warnings.filterwarnings("ignore", category=RuntimeWarning)


def golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle (builds oracle/_build/liboracle.so on first use)."""
    import subprocess

    if not os.path.isfile(os.path.join(ROOT, "oracle", "_build", "liboracle.so")):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "_build/liboracle.so"])
    from oracle import oracle as orc

    return orc


def flat_result(res):
    return np.concatenate([
        np.array([np.nan if v is None else v for v in np.atleast_1d(np.array(r, dtype=object)).tolist()], dtype=float)
        for r in res
    ])


# (tag, function, SPARCParams kwargs, LDPCParams args, extra kwargs, repetitions) -- mirrors
# tests/golden/gen_golden.py::gen_flows
FLOW_CASES = [
    ("plain_c1", "amp_ldpc_sim", dict(L=128, M=4, sigma=0.708, p=2, r=1, t=64), None, {}, 4),
    ("orig_s", "amp_ldpc_sim", dict(L=64, M=8, sigma=0.8, p=4, r=1, t=64), ("802.16", "5/6", 4), {}, 4),
    ("soft_s", "soft_amp_ldpc_sim", dict(L=64, M=8, sigma=0.8, p=4, r=1, t=64), ("802.16", "5/6", 8), dict(soft_iter=2), 4),
    ("hard_s", "hardinitbeta_amp_ldpc_sim", dict(L=64, M=8, sigma=0.8, p=4, r=1, t=64), ("802.16", "5/6", 8), {}, 4),
    ("thr_s", "soft_amp_ldpc_hardinit", dict(L=64, M=8, sigma=0.85, p=4, r=1, t=64), ("802.16", "5/6", 8),
     dict(soft_iter=3, threshold=0.6), 4),
    ("thr_m32", "soft_amp_ldpc_hardinit", dict(L=96, M=32, sigma=1.0, p=4, r=1, t=64), ("802.16", "1/2", 20),
     dict(soft_iter=3, threshold=0.7), 3),
    ("soft_m32", "soft_amp_ldpc_sim", dict(L=96, M=32, sigma=1.0, p=4, r=1, t=64), ("802.16", "1/2", 20),
     dict(soft_iter=2), 3),
]
