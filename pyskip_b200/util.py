"""Slice / shape helpers (src/pyskip/util.py:5-38)."""
from collections.abc import Iterable


def unify_slices(slices):
    """A bare index or slice becomes a 1-tuple; an all-int tuple stays a position; a mixed
    tuple has its ints widened to one-element slices (util.py:5-19)."""
    if not isinstance(slices, Iterable):
        return (slices,)
    items = tuple(slices)
    if all(isinstance(s, int) for s in items):
        return items
    out = []
    for s in items:
        if isinstance(s, int):
            out.append(slice(s, s + 1))
        elif isinstance(s, slice):
            out.append(s)
        else:
            raise AssertionError(f"unsupported index {s!r}")
    return tuple(out)


def broadcast_slice(ndim, sl, axis=None):
    """`sl` on `axis` (or on every axis when axis is None), full slices elsewhere
    (util.py:22-29)."""
    assert axis is None or 0 <= axis < ndim
    return tuple(sl if axis in (None, i) else slice(None) for i in range(ndim))


def broadcast_shape(ndim, shape):
    """An int or 1-tuple is repeated `ndim` times (util.py:32-38)."""
    if isinstance(shape, int):
        return (shape,) * ndim
    shape = tuple(shape)
    if len(shape) == 1:
        return shape * ndim
    assert len(shape) == ndim
    return shape
