"""``autograd.numpy`` stand-in: plain numpy.

The reference only applies ``autograd.numpy`` functions to concrete arrays
outside of any derivative trace (shape checks, norms, ``einsum`` on evaluated
derivative arrays), so plain numpy is a faithful substitute.  TEST
INFRASTRUCTURE ONLY - see ``oracle/refshim/autograd/__init__.py``.
"""
from numpy import *  # noqa: F401,F403
from numpy import linalg, random  # noqa: F401
