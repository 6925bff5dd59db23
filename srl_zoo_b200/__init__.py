"""srl_zoo_b200 -- B200-native train-step path for araffin/srl-zoo's conv AE / beta-VAE / DAE.

The hot path runs in libsrlz.so (hand-written sm_100a CUDA behind a C ABI, include/srlz.h), loaded on the first call
into it; a missing library raises there -- there is no PyTorch or CPU fallback for the hot path."""
from ._lib import LIB_PATH, lib  # noqa: F401  (lazy handle: raises on first use if the CUDA library is missing)
from .modules import B200SRLModules, B200SRLModulesSplit  # noqa: F401
from .engine import TrainStep  # noqa: F401
from . import losses, ops  # noqa: F401
from .install import install  # noqa: F401
