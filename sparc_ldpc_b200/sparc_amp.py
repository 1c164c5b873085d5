"""Drop-in for the functions of the reference's tutorial notebook `sparc_amp.ipynb` that its modules do not carry:
`pa_original` (cell 9), `pa_iterative` (cell 13), `bitwise_posterior` (cell 17) and the single-codeword simulation
`amp_sim` (cell 19; recorded run in cell 21).  `pa_parameterised`, `amp`, `sparc_transforms` and the fast transform are
the ones of `sparc_ldpc.py` (cells 4-6, 11, 15 are their ancestors).  Host logic stays on the host; the decode is one
libsparc_b200 AMP launch."""
import numpy as np

from . import sparc_ldpc as S


def pa_original(L, C, P, a=1.0):
    """Exponentially decaying allocation P_l ~ 2^(-2 a C l / L), scaled to sum P (cell 9; the notebook reads `a` from a
    global, here it is an argument with the classical value 1)."""
    pa = 2.0 ** (-2 * a * C * np.arange(L) / L)
    pa /= pa.sum() / P
    return pa


def pa_iterative(L, B, sigma, P, R_PA):
    """Iterative power allocation (cell 13): blocks of L // B sections get the power that makes them decodable at rate
    R_PA given the interference of the sections not yet allocated, until spreading the remaining power evenly gives more."""
    PA = np.zeros(L)
    tau = np.zeros(B)
    k = L // B
    for b in range(B):
        Premain = P - PA.sum()
        tau[b] = np.sqrt(sigma ** 2 + Premain)
        Pblock = 2 * np.log(2) * (R_PA / L) * tau[b] ** 2
        Pspread = Premain / (L - k * b)
        if Pblock > Pspread:
            PA[k * b:k * (b + 1)] = Pblock
        else:
            PA[k * b:] = Pspread
            break
    return PA


def bitwise_posterior(beta, L, M):
    """Bit-1 posteriors of every section, each section normalised by its own sum (cell 17; known answer in cell 25)."""
    b = np.asarray(beta, dtype=np.float64).reshape(L, M)
    assert M % 2 == 0
    return S.sp2bp((b / b.sum(axis=1, keepdims=True)).reshape(-1), L, M)


def amp_sim(L, M, sigma_n, P, R, T, R_PA, full=False, mode=None):
    """One plain SPARC codeword through the AMP decoder (cell 19), with the notebook's draw order on the global numpy
    stream (randint(0, M, L) for the message, then randn(n, 1) for the noise) and its seed-0 design matrix.  Returns the
    section error rate `1 - correct` like the notebook; full=True returns the record the notebook's run in cell 21 shows
    ({'C', 'EbN0', 'L', 'M', 'P', 'R', 'R_PA', 'T', 'ber', 'fc', 'n', 'ser', 'sigma_n', 'snr'})."""
    snr = P / sigma_n ** 2
    C = 0.5 * np.log2(1 + snr)
    n = int(L * np.log2(M) / R)
    Pl = pa_iterative(L, L, sigma_n, P, R_PA)
    tx_message = np.random.randint(0, M, L).tolist()
    Ab, Az, _ = S.sparc_transforms(L, M, n)
    beta_0 = np.zeros((L * M, 1))
    beta_0[np.arange(L) * M + np.asarray(tx_message)] = np.sqrt(n * Pl).reshape(-1, 1)
    x = Ab(beta_0)
    z = np.random.randn(n, 1) * sigma_n
    y = (x + z).reshape(-1, 1)
    from . import engine as E
    prev = E.AMP_MODE
    if mode is not None:
        E.AMP_MODE = mode
    try:
        beta = S.amp(y, sigma_n, Pl, L, M, T, Ab, Az).reshape(-1)
    finally:
        E.AMP_MODE = prev
    rx_message = beta.reshape(L, M).argmax(axis=1).tolist()
    correct = float(np.sum(np.array(rx_message) == np.array(tx_message)) / L)
    if not full:
        return 1 - correct
    ber = float(sum(bin(a ^ b).count("1") for a, b in zip(tx_message, rx_message)) / (L * np.log2(M)))
    return {"L": L, "M": M, "sigma_n": sigma_n, "P": P, "R": R, "T": T, "snr": snr, "C": float(C), "n": n, "fc": correct,
            "R_PA": R_PA, "ber": ber, "EbN0": 1 / (2 * R) * (P / sigma_n ** 2), "ser": 1 - correct}
