"""Stand-in for matplotlib's `pylab` (not installed in this image).

TEST INFRASTRUCTURE ONLY.  `from pylab import *` in the reference
(ldpc/sparc_ldpc.py:3, ldpc/amp_exit.py:4, ldpc/amp_test.py:3) is used for its
numpy star-exports: bare `sum`, `log`, `log2`, `linspace` are numpy's.  2018-era
pylab did NOT shadow the builtins max/min/abs/round/pow/bool (numpy 2 exports
them), and the reference relies on the builtin `max` (sparc_ldpc.py:54,110).
"""
import builtins as _b
from numpy import *            # noqa: F401,F403
from numpy.fft import *        # noqa: F401,F403
from numpy.random import *     # noqa: F401,F403
from numpy.linalg import *     # noqa: F401,F403
import numpy as np             # noqa: F401

max = _b.max
min = _b.min
abs = _b.abs
round = _b.round
pow = _b.pow
bool = _b.bool
bytes = _b.bytes
int = _b.int
float = _b.float
complex = _b.complex
str = _b.str
object = _b.object
