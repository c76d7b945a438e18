"""Reductions (src/pyskip/reduce.py:8-50).

``reduce(op, t, axis)`` is the reference's pairwise tree: odd extent -> op(reduce(t[:-1]), t[-1]);
even extent -> reduce(op(t[0::2], t[1::2])) -- log2(n) strided two-source evaluations.  For the
built-in full reductions (``sum``/``add``, ``prod``/``mul``, ``maximum``, ``minimum`` with
axis=None) the GPU backend does one pass over the runs instead (sum(val * len)): for integers
the result is identical (arithmetic mod 2^32 is order independent); for float32 it accumulates
in fp64 and rounds once, which agrees with the reference's fp32 pairwise tree to ~1e-7
relative (bound used by the tests: 1e-6).  The same holds per row for ``axis=0`` (the
fastest axis): one segmented pass (psk_reduce_segments) instead of the tree."""
import operator

from .index import take
from .manipulate import flatten, squeeze
from .tensor import Tensor

_SUM, _MAX, _MIN, _PROD = 0, 1, 2, 3


def reduce(op, t, axis=None, keepdims=False):
    assert len(t) > 0
    if axis is None:
        return reduce(op, flatten(t), axis=0, keepdims=keepdims)
    assert 0 <= axis < len(t.shape)
    if len(t) == 1:
        return t if keepdims else t.item()
    n = t.shape[axis]
    if n == 1:
        return t if keepdims else squeeze(t)
    if n % 2 == 1:
        last = take(t, slice(-1, n), axis)
        if len(t.shape) == 1:
            last = last.item()
        return op(reduce(op, take(t, slice(-1), axis)), last)
    lo = take(t, slice(0, n, 2), axis)
    hi = take(t, slice(1, n, 2), axis)
    return reduce(op, op(lo, hi), axis, keepdims)


def _axis0(t, kind, keepdims):
    """Reduction of the fastest axis in one pass over the runs (psk_reduce_segments): segment j
    is the row of t.shape[0] consecutive flat elements at outer coordinate j."""
    if t.dtype not in (int, float) or len(t.shape) < 2:
        return None
    flat = t.to_array()._reduce_segments(kind, t.shape[0])
    shape = ((1,) + tuple(t.shape[1:])) if keepdims else tuple(t.shape[1:])
    return Tensor(shape=shape, val=flat, dtype=t.dtype)


def _builtin(t, kind, axis, keepdims, op):
    # The reference's tree reduces an axis of an N-D tensor correctly only when its extent is a
    # power of two (an odd extent recurses WITHOUT the axis, reduce.py:24-29, and folds the whole
    # tensor into the partial result); parity with the reference comes first, so the one-pass
    # kernel is used exactly where the tree computes the per-row reduction.
    n = t.shape[axis] if 0 <= axis < len(t.shape) else 0
    if axis == 0 and n > 0 and n & (n - 1) == 0:
        out = _axis0(t, kind, keepdims)
        if out is not None:
            return out
    return reduce(op, t, axis, keepdims)


def _full(t, kind, keepdims):
    value = t.to_array()._reduce(kind)
    if keepdims:
        return Tensor(shape=(1,), val=value, dtype=t.dtype)
    return value


def add(t, axis=None, keepdims=False):
    if t.empty():
        return t.dtype(0)
    if axis is None:
        return _full(t, _SUM, keepdims)
    return _builtin(t, _SUM, axis, keepdims, operator.add)


def mul(t, axis=None, keepdims=False):
    if t.empty():
        return t.dtype(1)
    if axis is None:
        return _full(t, _PROD, keepdims)
    return _builtin(t, _PROD, axis, keepdims, operator.mul)


def maximum(t, axis=None, keepdims=False):
    if axis is None:
        return _full(t, _MAX, keepdims)
    return _builtin(t, _MAX, axis, keepdims, lambda a, b: a.max(b))


def minimum(t, axis=None, keepdims=False):
    if axis is None:
        return _full(t, _MIN, keepdims)
    return _builtin(t, _MIN, axis, keepdims, lambda a, b: a.min(b))


sum = add
prod = mul
