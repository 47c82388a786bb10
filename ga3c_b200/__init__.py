"""ga3c_b200 -- B200-native implementation of GA3C's predict / train hot path.

Host side mirrors the reference's interface for this path (same names, argument meaning, error
behaviour): `Network` (NetworkVP.py), `ThreadPredictor`, `ThreadTrainer`, `Config`.  All device work
goes through the C-ABI library `ga3c_b200/_lib/libga3c_b200.so` (include/ga3c_b200.h); there is no
CPU fallback -- importing works anywhere, but constructing a `Network` without the library or
without an sm_100 GPU raises.
"""
from .config import Config  # noqa: F401
from .network import Network  # noqa: F401
from .threads import LockstepTrainer, ThreadPredictor, ThreadTrainer  # noqa: F401
from .batcher import NativePredictor  # noqa: F401

__all__ = ["Config", "Network", "ThreadPredictor", "ThreadTrainer", "LockstepTrainer", "NativePredictor"]
