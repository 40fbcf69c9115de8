"""pynndescent_b200 -- B200 (sm_100a) implementation of PyNNDescent's NN-descent hot path.

Importing this package loads libnnd_b200.so; it raises if the CUDA library is absent
(no CPU fallback by design).
"""
from . import _ffi  # noqa: F401  (fails loudly if the CUDA library is missing)
from .pynndescent_ import NNDescent  # noqa: F401
from .transformer import PyNNDescentTransformer  # noqa: F401
from . import distances  # noqa: F401

__version__ = "0.1.0"
