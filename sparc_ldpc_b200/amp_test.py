"""Reference-facing mirror of ldpc/amp_test.py: `amp_test` is `amp` that also returns the iteration index
at which it stopped (amp_test.py:14-50); `amp_init_test` compares a warm start with a cold start
(amp_test.py:53-110)."""
import numpy as np

from . import sparc_ldpc as S
from .sparc_ldpc import bits2indices, sparc_transforms


def amp_test(y, sigma_n, Pl, L, M, T, Ab, Az, beta=None):
    """-> (beta (LM,1), t): t is the loop index at which tau == last_tau fired, else T-1."""
    return S._amp_host(y, Pl, L, M, T, Ab, Az, beta)


def amp_init_test(L, M, snr_dB, P, r_sparc):
    """amp_test.py:53-110 -> (ber_init [1], ber_no_init [1])."""
    logm = np.log2(M)
    total_bits = int(L * logm)
    sigma = np.sqrt(P / 10 ** (snr_dB / 20))
    n = int(L * np.log2(M) / r_sparc)
    Pl = P / L * np.ones(L)
    idx = bits2indices(np.random.randint(0, 2, total_bits).tolist(), M)
    Ab, Az, _ = sparc_transforms(L, M, n)
    beta = np.zeros((L * M, 1))
    for l in range(L):
        beta[l * M + idx[l]] = np.sqrt(n * Pl[l])
    y = (Ab(beta) + np.random.randn(n, 1) * sigma).reshape(-1, 1)
    b_init, t_init = amp_test(y, 0, Pl, L, M, 64, Ab, Az, beta)
    b_none, t_none = amp_test(y, 0, Pl, L, M, 64, Ab, Az)
    rx_i = np.argmax(b_init.reshape(L, M), axis=1)
    rx_n = np.argmax(b_none.reshape(L, M), axis=1)
    cnt = lambda rx: sum(bin(int(a) ^ int(b)).count("1") for a, b in zip(idx, rx)) / total_bits
    print("For initialised amp, BER= ", [cnt(rx_i)], " and iterations= ", t_init)
    print("For amp with all zero beta_0, BER= ", [cnt(rx_n)], " and iterations= ", t_none)
    return [cnt(rx_i)], [cnt(rx_n)]
