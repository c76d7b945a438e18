"""`Tensor`: the user-facing N-D array (1 to 3 dimensions, as in the reference) backed by a
lazy run-length-encoded array on the GPU.

Mirrors src/pyskip/tensor.py:22-350: a tensor wraps one of the extension's
``Tensor{1,2,3}{i,f,b}`` objects; every operator forwards to the flat array of that object and
re-wraps the result with the same shape (tensor.py:131-157).  Index order is (x, y, z) with x
fastest; numpy shapes are reversed on the way in and out (tensor.py:78-81, 323-325).
"""
import functools
import operator
import re
from collections.abc import Iterable

import numpy as np

from . import _pyskip_b200_ext as _ext
from .exceptions import (IncompatibleTensorError, InvalidTensorError, TypeConversionError,
                         UnimplementedOperationError)
from .util import unify_slices

_LETTER = {int: "i", float: "f", bool: "b"}
_PREFIX = {int: "Int", float: "Float", bool: "Bool"}
_DTYPE_OF_LETTER = {v: k for k, v in _LETTER.items()}
_DTYPE_OF_PREFIX = {v: k for k, v in _PREFIX.items()}


def _tensor_class(ndim, dtype):
    try:
        return getattr(_ext, f"Tensor{ndim}{_LETTER[dtype]}")
    except KeyError:
        raise TypeConversionError(f"unsupported dtype {dtype!r}") from None


def _array_dtype(cpp_array):
    m = re.match(r"(Int|Float|Bool)Array$", type(cpp_array).__name__)
    return _DTYPE_OF_PREFIX[m.group(1)] if m else None


class TensorBuilder:
    """Point-wise construction (tensor.py:22-54): writes go to a host-side builder and the
    tensor is uploaded once by build()."""

    def __init__(self, shape, val, dtype):
        if len(shape) > 3:
            raise UnimplementedOperationError("Builders only support up to three dimensions")
        self.shape = tuple(shape)
        self.val = val
        self.dtype = dtype
        size = functools.reduce(operator.mul, self.shape)
        self._builder = getattr(_ext, f"{_PREFIX[dtype]}Builder")(size, val)
        self._scale = [1]
        for extent in self.shape[:-1]:
            self._scale.append(self._scale[-1] * extent)

    def __setitem__(self, items, value):
        if not isinstance(items, Iterable):
            items = (items,)
        elif not all(isinstance(i, int) for i in items):
            raise UnimplementedOperationError("Builder __setitem__ only works with integers for now")
        self._builder[sum(i * s for i, s in zip(items, self._scale))] = value

    def build(self):
        return Tensor(shape=self.shape, val=self._builder.build(), dtype=self.dtype)


class Tensor:
    # ---- construction -----------------------------------------------------------------
    def __init__(self, shape=None, val=0, dtype=int, cpp_tensor=None):
        if cpp_tensor is not None:
            self._adopt(cpp_tensor)
            return
        if shape is None:
            raise InvalidTensorError("Must specify shape")
        if isinstance(shape, int):
            shape = (shape,)
        shape = tuple(shape)
        if len(shape) < 1:
            raise InvalidTensorError("Dimensionality must be >= 1")
        if len(shape) > 3:
            raise InvalidTensorError("Only 1D, 2D and 3D tensors are supported")
        inferred = _array_dtype(val)
        if inferred is not None:
            dtype = inferred      # the array already carries its element type
        self._adopt(_tensor_class(len(shape), dtype)(shape, val))

    def _adopt(self, cpp_tensor):
        self._tensor = cpp_tensor
        m = re.search(r"Tensor(\d)(\w)$", type(cpp_tensor).__name__)
        self.ndim = int(m.group(1))
        self.dtype = _DTYPE_OF_LETTER[m.group(2)]
        self.shape = tuple(cpp_tensor.shape())
        return self

    @classmethod
    def builder(cls, shape, val=0, dtype=int):
        return TensorBuilder(shape, val, dtype)

    @classmethod
    def wrap(cls, cpp_tensor):
        return cls(cpp_tensor=cpp_tensor)

    @classmethod
    def force_tensor(cls, item, dtype):
        return item if isinstance(item, Tensor) else cls(shape=(1,), val=item, dtype=dtype)

    @classmethod
    def from_numpy(cls, np_arr):
        np_arr = np.asarray(np_arr)
        if np_arr.size == 0:
            raise InvalidTensorError("cannot build a tensor from an empty array")
        if np_arr.dtype == np.float64:
            np_arr = np_arr.astype(np.float32)
        elif np_arr.dtype.kind in "iu" and np_arr.dtype != np.int32:
            np_arr = np_arr.astype(np.int32)
        flat = _ext.from_numpy(np.ascontiguousarray(np_arr).reshape(-1))
        return cls(shape=tuple(reversed(np_arr.shape)), val=flat)

    @classmethod
    def from_list(cls, values, dtype=None):
        arr = np.array(values, dtype=dtype)
        if arr.size == 0:
            return _EmptyTensor(int if dtype is None else dtype)
        return cls.from_numpy(arr)

    # ---- operator forwarding (tensor.py:131-157) ------------------------------------------
    def _unary(self, method):
        out = getattr(self._tensor.array(), method)()
        return Tensor.wrap(_tensor_class(self.ndim, _array_dtype(out) or self.dtype)(self.shape, out))

    def _binary(self, other, method):
        if isinstance(other, Tensor):
            if self.shape != other.shape and other.shape != (1,) and self.shape != (1,):
                raise IncompatibleTensorError(f"Incompatible shapes: {self.shape} and {other.shape}")
            rhs = other._tensor.array()
        else:
            rhs = other
        out = getattr(self._tensor.array(), method)(rhs)
        if out is NotImplemented:
            return NotImplemented
        return Tensor.wrap(_tensor_class(self.ndim, _array_dtype(out) or self.dtype)(self.shape, out))

    def _inplace(self, other, method):
        return self._adopt(self._binary(other, method)._tensor)

    def abs(self):
        return self._unary("abs")

    def sqrt(self):
        return self._unary("sqrt")

    def exp(self):
        return self._unary("exp")

    def coalesce(self, other):
        return self._binary(other, "coalesce")

    def max(self, other):
        return self._binary(other, "max")

    def min(self, other):
        return self._binary(other, "min")

    __hash__ = None

    # ---- element access ---------------------------------------------------------------------
    def __setitem__(self, slices, value):
        slices = unify_slices(slices)
        self._tensor[slices] = value._tensor if isinstance(value, Tensor) else value

    def __getitem__(self, slices):
        out = self._tensor[unify_slices(slices)]
        return out if isinstance(out, self.dtype) else Tensor.wrap(out)

    def __len__(self):
        return len(self._tensor)

    def item(self):
        assert len(self) == 1
        return self._tensor.array()[0]

    # ---- conversions ------------------------------------------------------------------------
    def to(self, dtype):
        cast = getattr(self.to_array(), dtype.__name__)()
        return Tensor.wrap(_tensor_class(self.ndim, dtype)(self.shape, cast))

    def to_array(self):
        return self._tensor.array()

    def to_numpy(self):
        return self._tensor.array().to_numpy().reshape(tuple(reversed(self.shape)))

    def to_list(self):
        return self.to_numpy().tolist()

    def to_runs(self):
        return self.to_array().runs()

    def rle_length(self):
        return len(self.to_runs()[0])

    def to_string(self, threshold=20, separator=" "):
        from .io import format_tensor
        return format_tensor(self, threshold, separator)

    def __str__(self):
        return self.to_string()

    def __repr__(self):
        head = f"Tensor(shape={self.shape}, dtype={self.dtype.__name__})"
        body = ("\n" + self.to_string(separator=", ")).replace("\n", "\n    ")
        return f"{head}:{body}"

    # ---- misc ----------------------------------------------------------------------------------
    def empty(self):
        return len(self) == 0

    def clone(self):
        return self._unary("clone")

    def eval(self):
        return self._unary("eval")

    def reshape(self, shape):
        from .manipulate import reshape
        return reshape(self, shape)

    def flatten(self, shape=None):
        from .manipulate import flatten
        return flatten(self)


class _EmptyTensor(Tensor):
    """`Tensor.from_list([])`: length 0, only good for `empty()` / `len()` / reductions
    (src/pyskip/tests/reduce_test.py:23)."""

    def __init__(self, dtype):
        self.dtype = dtype
        self.ndim = 1
        self.shape = (0,)
        self._tensor = None

    def __len__(self):
        return 0


def _install_operators():
    unary = {"__neg__": "__neg__", "__pos__": "__pos__", "__abs__": "__abs__", "__invert__": "__invert__"}
    for name, method in unary.items():
        setattr(Tensor, name, functools.partialmethod(Tensor._unary, method))
    binary = ["add", "sub", "mul", "truediv", "floordiv", "mod", "pow", "and", "or", "xor", "lshift", "rshift"]
    for name in binary:
        setattr(Tensor, f"__{name}__", functools.partialmethod(Tensor._binary, method=f"__{name}__"))
        setattr(Tensor, f"__r{name}__", functools.partialmethod(Tensor._binary, method=f"__r{name}__"))
        setattr(Tensor, f"__i{name}__", functools.partialmethod(Tensor._inplace, method=f"__{name}__"))
    for name in ["eq", "ne", "lt", "le", "gt", "ge"]:
        setattr(Tensor, f"__{name}__", functools.partialmethod(Tensor._binary, method=f"__{name}__"))


_install_operators()
