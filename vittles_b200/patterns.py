"""Flatten / fold patterns with the interface of ``paragami``.

The reference's users - and its own tests (``vittles/tests/test_utils.py:23-75``,
``test_sparse_hessian_lib.py:15-40``, ``test_lr_cov_lib.py:20-61``) - describe
structured, constrained parameters with ``paragami`` patterns and hand vittles the
*flattened* objective (``paragami.FlattenFunctionInput``).  ``paragami`` is built on
``autograd.numpy``; this module offers the same classes and call signatures on
``torch`` so that such code can target :mod:`vittles_b200`, whose objectives are torch
callables: ``fold`` is differentiable by ``torch.func`` and runs on whatever device the
flat vector lives on, ``flatten`` accepts numpy arrays or tensors.

    NumericArrayPattern(shape, lb, ub)   NumericVectorPattern(length, lb, ub)
    PSDSymmetricMatrixPattern(size)      PatternDict()        PatternArray(array_shape, base_pattern)
    FlattenFunctionInput(fun, patterns, free, argnums)
    FlattenFunctionInputAndOutput(fun, input_patterns, input_free, input_argnums,
                                  output_patterns, output_free, output_retnums)

Free (unconstrained) maps, as in paragami: one-sided bound ``log(x - lb)`` / ``log(ub - x)``,
two-sided ``logit((x - lb) / (ub - lb))``, unbounded identity; a PSD matrix is the vector of
the lower triangle of its Cholesky factor with a log diagonal (row-major order of the
triangle); the non-free flat form is the plain raveled array.  ``paragami`` itself is not
available in this build's image, so these maps are restated from its documentation, and
``tests/test_gpu_patterns.py`` pins them against the closed forms of the reference's
QuadraticModel fixture (every truth there is derived from the same maps).
"""
from collections import OrderedDict

import numpy as np
import torch


def _t(x, like=None):
    if isinstance(x, torch.Tensor):
        return x
    dev = like.device if isinstance(like, torch.Tensor) else None
    return torch.as_tensor(np.asarray(x, dtype=np.float64), device=dev)


def _ret(t, proto):
    """numpy in -> numpy out; tensor in -> tensor out"""
    return t if isinstance(proto, torch.Tensor) else t.detach().cpu().numpy()


class Pattern:
    def flat_length(self, free):
        raise NotImplementedError

    def empty(self, valid=True):
        raise NotImplementedError

    def random(self):
        return self.fold(np.random.normal(size=self.flat_length(True)), free=True)

    def empty_bool(self, value):
        return np.full(self.flat_length(False), bool(value))


class NumericArrayPattern(Pattern):
    def __init__(self, shape, lb=-float('inf'), ub=float('inf'), default_validate=True):
        self._shape = tuple(int(s) for s in shape)
        self._lb, self._ub = float(lb), float(ub)
        if not self._lb < self._ub:
            raise ValueError('Upper bound ub must strictly exceed lower bound lb')
        self._size = int(np.prod(self._shape)) if self._shape else 1

    def shape(self):
        return self._shape

    def bounds(self):
        return self._lb, self._ub

    def flat_length(self, free=None):
        return self._size

    def empty(self, valid=True):
        if not valid:
            return np.empty(self._shape)
        lb, ub = self._lb, self._ub
        v = 0.0 if (lb < 0 < ub) else (lb + 1.0 if np.isfinite(lb) else ub - 1.0)
        if np.isfinite(lb) and np.isfinite(ub):
            v = 0.5 * (lb + ub)
        return np.full(self._shape, v)

    def validate_folded(self, folded_val, validate_value=None):
        f = _t(folded_val)
        if tuple(f.shape) != self._shape:
            return False, 'Wrong size: {} instead of {}'.format(tuple(f.shape), self._shape)
        if bool((f < self._lb).any()):
            return False, 'Value beneath lower bound.'
        if bool((f > self._ub).any()):
            return False, 'Value above upper bound.'
        return True, ''

    def flatten(self, folded_val, free, validate_value=None):
        x = _t(folded_val)
        try:
            ok, msg = self.validate_folded(x)
        except RuntimeError:                      # inside a torch.func transform: values are not inspectable
            ok, msg = True, ''
        if not ok:
            raise ValueError(msg)
        x = x.reshape(-1)
        if free:
            lb, ub = self._lb, self._ub
            if np.isfinite(lb) and np.isfinite(ub):
                u = (x - lb) / (ub - lb)
                x = torch.log(u) - torch.log1p(-u)
            elif np.isfinite(lb):
                x = torch.log(x - lb)
            elif np.isfinite(ub):
                x = torch.log(ub - x)
        return _ret(x, folded_val)

    def fold(self, flat_val, free, validate_value=None):
        x = _t(flat_val)
        if x.dim() != 1 or x.numel() != self._size:
            raise ValueError('Wrong length for array flat value: {} instead of {}'.format(tuple(x.shape), self._size))
        if free:
            lb, ub = self._lb, self._ub
            if np.isfinite(lb) and np.isfinite(ub):
                x = lb + (ub - lb) * torch.sigmoid(x)
            elif np.isfinite(lb):
                x = torch.exp(x) + lb
            elif np.isfinite(ub):
                x = ub - torch.exp(x)
        return _ret(x.reshape(self._shape), flat_val)


class NumericVectorPattern(NumericArrayPattern):
    def __init__(self, length, lb=-float('inf'), ub=float('inf'), default_validate=True):
        super().__init__((int(length),), lb=lb, ub=ub, default_validate=default_validate)


class NumericScalarPattern(NumericArrayPattern):
    def __init__(self, lb=-float('inf'), ub=float('inf'), default_validate=True):
        super().__init__((1,), lb=lb, ub=ub, default_validate=default_validate)


class PSDSymmetricMatrixPattern(Pattern):
    """Symmetric positive-definite ``size x size`` matrices; the diagonal may be bounded below (``diag_lb``)."""

    def __init__(self, size, diag_lb=0.0, default_validate=True):
        self._n = int(size)
        self._diag_lb = float(diag_lb)
        if self._diag_lb < 0:
            raise ValueError('The diagonal lower bound diag_lb must be >-= 0.')
        r, c = np.tril_indices(self._n)
        self._rows, self._cols = torch.as_tensor(r), torch.as_tensor(c)
        self._is_diag = torch.as_tensor(r == c)

    def size(self):
        return self._n

    def shape(self):
        return (self._n, self._n)

    def flat_length(self, free):
        return self._n * (self._n + 1) // 2 if free else self._n * self._n

    def empty(self, valid=True):
        return np.eye(self._n) * (self._diag_lb + 1.0) if valid else np.empty((self._n, self._n))

    def flatten(self, folded_val, free, validate_value=None):
        a = _t(folded_val)
        if tuple(a.shape) != (self._n, self._n):
            raise ValueError('The matrix is not of shape ({}, {})'.format(self._n, self._n))
        if not free:
            return _ret(a.reshape(-1), folded_val)
        a = a - self._diag_lb * torch.eye(self._n, dtype=a.dtype, device=a.device)
        L = torch.linalg.cholesky(a)
        v = L[self._rows.to(a.device), self._cols.to(a.device)]
        v = torch.where(self._is_diag.to(a.device), torch.log(v), v)
        return _ret(v, folded_val)

    def fold(self, flat_val, free, validate_value=None):
        x = _t(flat_val)
        if x.dim() != 1 or x.numel() != self.flat_length(free):
            raise ValueError('Wrong length for PSDSymmetricMatrix flat value.')
        if not free:
            return _ret(x.reshape(self._n, self._n), flat_val)
        v = torch.where(self._is_diag.to(x.device), torch.exp(x), x)
        L = torch.zeros((self._n, self._n), dtype=x.dtype, device=x.device)
        L = L.index_put((self._rows.to(x.device), self._cols.to(x.device)), v)
        out = L @ L.T + self._diag_lb * torch.eye(self._n, dtype=x.dtype, device=x.device)
        return _ret(out, flat_val)


class PatternDict(Pattern):
    """Ordered dictionary of patterns; the flat vector is the concatenation in insertion order."""

    def __init__(self, free_default=None, default_validate=True):
        self._patterns = OrderedDict()
        self._locked = False

    def __setitem__(self, name, pattern):
        if self._locked:
            raise ValueError('The dictionary is locked, and its values cannot be changed.')
        self._patterns[name] = pattern

    def __getitem__(self, name):
        return self._patterns[name]

    def __delitem__(self, name):
        if self._locked:
            raise ValueError('The dictionary is locked, and its values cannot be changed.')
        del self._patterns[name]

    def keys(self):
        return self._patterns.keys()

    def lock(self):
        self._locked = True

    def flat_length(self, free):
        return sum(p.flat_length(free) for p in self._patterns.values())

    def empty(self, valid=True):
        return OrderedDict((k, p.empty(valid)) for k, p in self._patterns.items())

    def flatten(self, folded_val, free, validate_value=None):
        parts = [_t(p.flatten(folded_val[k], free)) for k, p in self._patterns.items()]
        proto = next(iter(folded_val.values())) if len(folded_val) else None
        dev = next((q.device for q in parts if q.is_cuda), None)
        out = torch.cat([q.to(dev) if dev is not None else q for q in parts]) if parts else torch.zeros(0, dtype=torch.float64)
        return _ret(out, proto) if proto is not None else out.numpy()

    def fold(self, flat_val, free, validate_value=None):
        x = _t(flat_val)
        if x.numel() != self.flat_length(free):
            raise ValueError('Wrong size for pattern dictionary: {} instead of {}'.format(x.numel(), self.flat_length(free)))
        out, off = OrderedDict(), 0
        for k, p in self._patterns.items():
            n = p.flat_length(free)
            out[k] = _ret(_t(p.fold(x[off:off + n], free)), flat_val)
            off += n
        return out

    def flat_indices(self, folded_bool, free):
        """Indices into the flat vector of the entries selected by a folded boolean structure (non-free layout, or
        free layout for patterns whose free and non-free lengths coincide)."""
        idx, off = [], 0
        for k, p in self._patterns.items():
            n = p.flat_length(free)
            mask = np.asarray(folded_bool[k]).reshape(-1)
            if mask.size != n:
                raise NotImplementedError('flat_indices of a pattern whose free form is not entrywise')
            idx.append(off + np.nonzero(mask)[0])
            off += n
        return np.concatenate(idx) if idx else np.zeros(0, dtype=int)


class PatternArray(Pattern):
    """An ``array_shape`` array of identical patterns; folded values have shape ``array_shape + base shape``."""

    def __init__(self, array_shape, base_pattern, default_validate=True):
        self._array_shape = tuple(int(s) for s in array_shape)
        self._base = base_pattern
        self._count = int(np.prod(self._array_shape)) if self._array_shape else 1
        self._base_shape = tuple(base_pattern.shape())

    def shape(self):
        return self._array_shape + self._base_shape

    def array_shape(self):
        return self._array_shape

    def base_pattern(self):
        return self._base

    def flat_length(self, free):
        return self._count * self._base.flat_length(free)

    def empty(self, valid=True):
        e = np.asarray(self._base.empty(valid))
        return np.broadcast_to(e, self._array_shape + e.shape).copy()

    def flatten(self, folded_val, free, validate_value=None):
        a = _t(folded_val)
        if tuple(a.shape) != self.shape():
            raise ValueError('Wrong shape for PatternArray: {} instead of {}'.format(tuple(a.shape), self.shape()))
        items = a.reshape((self._count,) + self._base_shape)
        out = torch.cat([_t(self._base.flatten(items[i], free)) for i in range(self._count)])
        return _ret(out, folded_val)

    def fold(self, flat_val, free, validate_value=None):
        x = _t(flat_val)
        n = self._base.flat_length(free)
        if x.numel() != self._count * n:
            raise ValueError('Wrong size for PatternArray flat value.')
        items = torch.stack([_t(self._base.fold(x[i * n:(i + 1) * n], free)) for i in range(self._count)])
        return _ret(items.reshape(self.shape()), flat_val)

    def flat_indices(self, folded_bool, free):
        mask = np.asarray(folded_bool).reshape(-1)
        if mask.size != self.flat_length(free):
            raise NotImplementedError('flat_indices of a pattern whose free form is not entrywise')
        return np.nonzero(mask)[0]


def _as_list(x, n=None):
    if isinstance(x, (list, tuple)):
        return list(x)
    return [x] if n is None else [x] * n


class FlattenFunctionInput:
    """``fun`` with the arguments ``argnums`` replaced by their flat vectors (``paragami.FlattenFunctionInput``):
    the returned callable folds those arguments with ``patterns`` (free or not) and calls ``fun``."""

    def __init__(self, original_fun, patterns, free, argnums=None):
        self._fun = original_fun
        self._patterns = _as_list(patterns)
        n = len(self._patterns)
        self._free = [bool(f) for f in _as_list(free, n)] if isinstance(free, (list, tuple)) else [bool(free)] * n
        self._argnums = list(range(n)) if argnums is None else [int(a) for a in _as_list(argnums)]
        if not (len(self._free) == len(self._argnums) == n):
            raise ValueError('patterns, free and argnums must have the same length')
        if len(set(self._argnums)) != n:
            raise ValueError('argnums must be unique')

    def _call(self, args, kwargs):
        """(result, numpy_in): the folded arguments are always handed to ``fun`` as tensors (objectives are torch
        callables); numpy_in tells whether every flat argument came in as numpy."""
        args = list(args)
        numpy_in = not any(isinstance(args[a], torch.Tensor) for a in self._argnums)
        for p, f, a in zip(self._patterns, self._free, self._argnums):
            args[a] = p.fold(_t(args[a]), free=f)
        return self._fun(*args, **kwargs), numpy_in

    def __call__(self, *args, **kwargs):
        ret, numpy_in = self._call(args, kwargs)
        if numpy_in and isinstance(ret, torch.Tensor):
            return ret.detach().cpu().numpy()
        return ret


class FlattenFunctionInputAndOutput:
    """Flat inputs as :class:`FlattenFunctionInput`, and the outputs ``output_retnums`` flattened with
    ``output_patterns`` (``paragami.FlattenFunctionInputAndOutput``)."""

    def __init__(self, original_fun, input_patterns, input_free, output_patterns, output_free, input_argnums=None,
                 output_retnums=None):
        self._flat_in = FlattenFunctionInput(original_fun, input_patterns, input_free, input_argnums)
        self._out_patterns = _as_list(output_patterns)
        n = len(self._out_patterns)
        self._out_free = [bool(f) for f in _as_list(output_free, n)] if isinstance(output_free, (list, tuple)) \
            else [bool(output_free)] * n
        self._retnums = list(range(n)) if output_retnums is None else [int(r) for r in _as_list(output_retnums)]

    def __call__(self, *args, **kwargs):
        ret, numpy_in = self._flat_in._call(args, kwargs)
        single = not isinstance(ret, tuple)
        vals = [ret] if single else list(ret)
        for p, f, r in zip(self._out_patterns, self._out_free, self._retnums):
            vals[r] = p.flatten(_t(vals[r]), free=f)
        if numpy_in:
            vals = [v.detach().cpu().numpy() if isinstance(v, torch.Tensor) else v for v in vals]
        return vals[0] if single else tuple(vals)
