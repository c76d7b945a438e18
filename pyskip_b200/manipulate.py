"""Shape manipulation (src/pyskip/manipulate.py:7-53): all of it is relabelling of the same
flat run-end array, or slice assignment into a fresh tensor."""
import functools
import operator

from .tensor import Tensor


def shape_len(shape):
    return functools.reduce(operator.mul, shape)


def reshape(t, shape):
    if isinstance(shape, int):
        shape = (shape,)
    shape = tuple(shape)
    assert len(t) == shape_len(shape)
    return Tensor(shape, val=t.to_array(), dtype=t.dtype)


def flatten(t):
    return reshape(t, len(t))


def squeeze(t, axis=None):
    kept = [s for i, s in enumerate(t.shape) if s != 1 or (axis is not None and axis != i)]
    return reshape(t, tuple(kept) if kept else (1,))


def stack(t1, t2):
    """New tensor whose shape is the element-wise sum of both shapes; the low corner holds
    t1, the high corner t2 (manipulate.py:30-41)."""
    assert t1.ndim == t2.ndim
    assert t1.dtype == t2.dtype
    out = Tensor(tuple(a + b for a, b in zip(t1.shape, t2.shape)), dtype=t1.dtype)
    out[tuple(slice(None, a) for a in t1.shape)] = t1
    out[tuple(slice(a, None) for a in t1.shape)] = t2
    return out


def concat(tensors):
    assert len(tensors) > 0
    if len(tensors) == 1:
        return tensors[0].clone()
    return stack(tensors[0], concat(tensors[1:]))
