"""Indexing helpers (src/pyskip/index.py:6-10)."""
from .util import broadcast_slice


def take(t, sl, axis=None):
    """Sub-tensor selected by `sl` along `axis` (every axis when axis is None)."""
    return t[broadcast_slice(t.ndim, sl, axis)]
