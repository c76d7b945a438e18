"""Configuration scopes (src/pyskip/config.py:8-58) over the extension's global config map.

Keys understood by the GPU backend: ``flush_tree_size_threshold`` (expression size above
which an array is evaluated eagerly; reference array.hpp:210) and ``gpu_eval_max_sources``
(sources merged per kernel launch, 2..30, default 8: wider expressions run as a chain of
launches because the k-way merge step costs O(k) per run).  ``accelerated_eval``,
``parallelize_threshold`` and ``parallelize_parts`` are accepted for compatibility and have
no effect (they select CPU code paths; results never depended on them).
``conv_stencil_kernel`` (bool, default True): run 3x3 / 3x3x3 convolutions on the one-pass
stencil-over-runs kernel instead of the tap loop (same results)."""
from contextlib import contextmanager

from . import _pyskip_b200_ext as _ext
from .exceptions import TypeConversionError

_cfg = _ext.config


def get_all_values():
    return _cfg.get_all_values()


def get_config_value(name, default=None):
    """Current value of a config key (or `default` when it is not set)."""
    return _cfg.get_all_values().get(name, default)


def set_value(name, val):
    if isinstance(val, bool):
        _cfg.set_bool_value(name, val)
    elif isinstance(val, int):
        _cfg.set_int_value(name, val)
    elif isinstance(val, float):
        _cfg.set_float_value(name, val)
    elif isinstance(val, str):
        _cfg.set_string_value(name, val)
    else:
        raise TypeConversionError(f"Unsupported type for config map: {type(val)}")


@contextmanager
def config_scope(**kwargs):
    saved = _cfg.get_all_values()
    try:
        for key, val in kwargs.items():
            set_value(key, val)
        yield
    finally:
        _cfg.set_all_values(saved)


@contextmanager
def num_threads_scope(num):
    with config_scope(parallelize_parts=int(num)):
        yield


@contextmanager
def flush_threshold_scope(num):
    with config_scope(flush_tree_size_threshold=int(num)):
        yield


@contextmanager
def lazy_evaluation_scope():
    with flush_threshold_scope(2 ** 30):
        yield


@contextmanager
def greedy_evaluation_scope():
    with flush_threshold_scope(0):
        yield
