"""``paragami`` stand-in.  ``sensitivity_lib.py:14`` imports
``FlattenFunctionInput`` but never uses it inside the library.  TEST
INFRASTRUCTURE ONLY."""


class FlattenFunctionInput:  # pragma: no cover - never instantiated
    def __init__(self, *a, **k):
        raise NotImplementedError('paragami is not available in this image')
