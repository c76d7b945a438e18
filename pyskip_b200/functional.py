"""`pyskip.functional` namespace (src/pyskip/functional.py:1-2)."""
from .convolve import *  # noqa: F401,F403
from .reduce import *  # noqa: F401,F403
