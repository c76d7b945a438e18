"""Exception types of the Python layer (src/pyskip/exceptions.py:1-22)."""


class PySkipError(RuntimeError):
    """Base class of every error raised by the Python layer."""


class PySkipInvariantViolation(PySkipError):
    """An internal invariant does not hold."""


class InvalidTensorError(PySkipError):
    """A tensor cannot be constructed from the given arguments."""


class IncompatibleTensorError(PySkipError):
    """Operands of an operation have incompatible shapes."""


class UnimplementedOperationError(PySkipError):
    """The operation is not supported (yet)."""


class TypeConversionError(PySkipError):
    """A Python value cannot be converted to a supported element type."""
