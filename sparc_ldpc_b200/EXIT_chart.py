"""Closed-form EXIT functions of the reference's ldpc/EXIT_chart.py (J / J_inverse :24-38, I_E_VND :40-43,
I_A_CND :45-47, I_E_REP :50-52) and the combined AMP+VND curves of ldpc/amp_exit.py:417-449.  Scalar host
formulas -- microseconds of CPU; kept so that a user of the reference's EXIT tooling finds the same names."""
import numpy as np

from .amp_exit import J  # same approximation as EXIT_chart.py:31-38


def J_inverse(I):
    """EXIT_chart.py:24-29 (unlike amp_exit.J_inverse this one does not clip I = 1)."""
    assert 0 <= I <= 1
    if I <= 0.3646:
        return 1.09542 * (I ** 2) + 0.214217 * I + 2.33727 * np.sqrt(I)
    return -0.706692 * np.log(0.386013 * (1 - I)) + 1.75017 * I


def I_E_VND(I_A, dv, EbN0, R):
    """Variable-node EXIT curve over BPSK/AWGN, EbN0 linear (EXIT_chart.py:40-43)."""
    return J(np.sqrt((dv - 1) * J_inverse(I_A) ** 2 + 8 * R * EbN0))


def I_A_CND(I_E, dc):
    """Inverse check-node curve (EXIT_chart.py:45-47)."""
    return 1 - J(J_inverse(1 - I_E) / np.sqrt(dc - 1))


def I_E_REP(I_A, dc):
    """I_E_CND = 1 - I_E_REP (EXIT_chart.py:49-52)."""
    return J(np.sqrt(dc - 1) * J_inverse(1 - I_A))


def I_E_VND_amp(I_A_VND, d_v, poly_coeff):
    """Combined AMP + degree-d_v variable node (amp_exit.py:417-432; uses amp_exit's clipping J_inverse)."""
    from .amp_exit import J_inverse as Jinv
    I_A_amp = J(np.sqrt(d_v) * Jinv(I_A_VND))
    I_E_amp = np.sum(poly_coeff * np.array([1, I_A_amp, I_A_amp ** 2, I_A_amp ** 3]))
    I_E_amp = np.clip(I_E_amp, a_min=None, a_max=0.9999)
    return J(np.sqrt((d_v - 1) * (Jinv(I_A_VND)) ** 2 + (Jinv(I_E_amp)) ** 2))


def I_E_VND_amp_array(I_A_VND_array, a_v, b_v, poly_coeff):
    """Edge-perspective mixture over the variable-node degrees (amp_exit.py:434-449)."""
    out = np.zeros(np.shape(I_A_VND_array))
    for i in range(len(a_v)):
        if a_v[i] != 0:
            for j in range(len(out)):
                out[j] += b_v[i] * I_E_VND_amp(I_A_VND_array[j], i, poly_coeff)
    return out
