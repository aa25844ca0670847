"""mcpy_b200 -- B200-native (sm_100a) energy-and-move engine behind mcpy's plug-in API.

Only the hot path lives here: the calculator that
``GrandCanonicalEnsemble`` / ``BatchedReplicaExchange`` call, the free-volume
cells, and the sharded replica driver.  Everything is backed by
``libmcpy_b200.so`` (hand-written CUDA, C ABI in ``include/mcpy_b200.h``);
importing the package never touches the GPU, creating an engine does, and
there is no CPU fallback.
"""
from .engine import Engine, EngineError  # noqa: F401
from ._lib import pinned_copy, pinned_empty  # noqa: F401

__version__ = '0.1.0'
__all__ = ['Engine', 'EngineError']
