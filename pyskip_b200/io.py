"""Printing (src/pyskip/io.py:7-18)."""
import numpy as np

from .index import take
from .manipulate import stack


def format_tensor(tensor, threshold=20, separator=" "):
    """numpy-style text of the tensor; axes longer than `threshold` keep only their first and
    last `threshold` entries."""
    shown = tensor
    for axis, extent in enumerate(tensor.shape):
        if extent > threshold:
            head = take(shown, slice(threshold), axis=axis)
            tail = take(shown, slice(-threshold, None), axis=axis)
            shown = stack(head, tail)
    return np.array2string(shown.to_numpy(), threshold=threshold, separator=separator)
