"""Convolutions over run-length-encoded tensors (src/pyskip/convolve.py:5-59).

Semantics are the reference's: pad by slice-assigning into a filled tensor, then accumulate
``kernel[tap] * padded[shifted window]`` over the taps in z -> y -> x order (cross-correlation,
no flip, accumulation in the tensor's dtype).  What differs is how the expression reaches the
GPU: the reference flushes the growing sum every few taps (array.hpp:204-214) and re-evaluates
the padded tensor each time; here the padded tensor is evaluated once, the kernel values are
read back once, and the whole tap sum is built lazily and evaluated in as few launches as the
planner's width limit allows (27 taps of a 3x3x3 kernel = one k-way merge launch)."""
import numpy as np

from . import _pyskip_b200_ext as _ext
from .config import get_config_value, lazy_evaluation_scope
from .tensor import Tensor
from .util import broadcast_shape


def _pad(tensor, pads, fill):
    shape = tuple(s + 2 * p for s, p in zip(tensor.shape, pads))
    if all(p == 0 for p in pads):
        return tensor
    padded = Tensor(shape=shape, dtype=tensor.dtype, val=fill)
    padded[tuple(slice(p, s - p) for p, s in zip(pads, shape))] = tensor
    return padded.eval()


def pad_3d_tensor(tensor, padding, fill):
    assert tensor.ndim == 3
    return _pad(tensor, broadcast_shape(3, padding), fill)


def _conv_stencil(tensor, kernel, pads, fill):
    """One launch of the stencil-over-runs kernel (psk_conv_nd) when it handles this kernel
    shape; None otherwise.  Same values and run structure as the tap loop below."""
    if tensor.dtype not in (int, float) or kernel.dtype != tensor.dtype:
        return None
    out_shape = tuple(s + 2 * p - k + 1 for s, p, k in zip(tensor.shape, pads, kernel.shape))
    if any(o <= 0 for o in out_shape):
        return None
    weights = np.ascontiguousarray(kernel.to_numpy()).ravel()     # numpy C order = pyskip x fastest
    weights = weights.astype(np.int32 if tensor.dtype is int else np.float32)
    flat = _ext._conv_nd(tensor._tensor.array(), list(tensor.shape), list(kernel.shape), weights, list(pads),
                         tensor.dtype(fill))
    if flat is None:
        return None
    return Tensor(shape=out_shape, dtype=tensor.dtype, val=flat)


def _conv(tensor, kernel, padding, fill):
    nd = tensor.ndim
    pads = broadcast_shape(nd, padding)
    if get_config_value("conv_stencil_kernel", True):
        out = _conv_stencil(tensor, kernel, pads, fill)
        if out is not None:
            return out
    s = _pad(tensor, pads, fill)
    kshape = kernel.shape
    out_shape = tuple(a - k + 1 for a, k in zip(s.shape, kshape))
    weights = kernel.to_numpy()                     # numpy axes are reversed: weights[z, y, x]
    out = Tensor(shape=out_shape, dtype=tensor.dtype, val=0)
    with lazy_evaluation_scope():
        for tap in np.ndindex(*reversed(kshape)):   # z -> y -> x, x fastest
            pos = tuple(reversed(tap))              # (x, y, z)
            window = tuple(slice(o, a - k + o + 1) for o, a, k in zip(pos, s.shape, kshape))
            out += tensor.dtype(weights[tap]) * s[window]
        out = out.eval()
    return out


def conv_2d(tensor, kernel, padding=0, fill=0):
    assert tensor.ndim == 2
    assert kernel.ndim == 2
    return _conv(tensor, kernel, padding, fill)


def conv_3d(tensor, kernel, padding=0, fill=0):
    assert tensor.ndim == 3
    assert kernel.ndim == 3
    return _conv(tensor, kernel, padding, fill)
