"""``autograd.core`` stand-in (``sensitivity_lib.py:7``).  The reference uses
``primitive``/``defvjp``/``defjvp`` only inside ``get_linear_function``
(``sensitivity_lib.py:44-80``); the golden fixtures never differentiate
through that function, so inert decorators are sufficient.  TEST
INFRASTRUCTURE ONLY."""


def primitive(f):
    return f


def defvjp(fun, *vjps, **kwargs):
    return None


def defjvp(fun, *jvps, **kwargs):
    return None
