"""B200-native batched time-stepping contact hot path with Moby's interface.

Host-side mirror of the reference interface for this path; all compute goes through the C ABI in
include/b200moby.h (moby_b200/libb200moby.so).  There is no CPU fallback: the loader raises if the
CUDA library is missing.
"""
from .capi import lib, SceneDesc, Counters, B200MobyError  # noqa: F401
from .scenes import SceneBatch  # noqa: F401
from .simulator import TimeSteppingSimulator  # noqa: F401
