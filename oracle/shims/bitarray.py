"""Minimal stand-in for the `bitarray` package (not installed in this image).

TEST INFRASTRUCTURE ONLY.  The reference uses exactly two things
(ldpc/sparc_ldpc.py:301-311): construction from a '0'/'1' string, `invert()`,
and implicit conversion to a numpy array as the exponent of `**`, which with the
real package yields one element per bit.
"""
import numpy as _np


class bitarray:
    def __init__(self, s=""):
        self._b = _np.array([c == "1" for c in s], dtype=bool)

    def invert(self):
        self._b = ~self._b

    def __len__(self):
        return len(self._b)

    def __getitem__(self, i):
        return self._b[i]

    def __array__(self, dtype=None, copy=None):
        return self._b if dtype is None else self._b.astype(dtype)
