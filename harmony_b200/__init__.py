"""harmony_b200 — B200-native (sm_100a) implementation of the Harmony hot loop.

Host side mirrors the reference's R surface (RunHarmony / harmonize / the ``harmony`` module
class); all arithmetic on cell matrices runs in hand-written CUDA behind include/harmony_b200.h.
Importing this package does not load the CUDA library; constructing ``harmony`` does, and fails
loudly if it is missing (there is no CPU fallback).
"""
from .harmony_option import harmony_options  # noqa: F401
from .ui import RunHarmony, prepare_inputs  # noqa: F401
from .utils import HarmonyConvergencePlot, harmonize  # noqa: F401

__version__ = "0.1.0"


def __getattr__(name):
    if name == "harmony":
        from .harmony import harmony
        return harmony
    raise AttributeError(name)
