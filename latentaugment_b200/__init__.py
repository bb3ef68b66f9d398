"""latentaugment_b200 -- B200-native LatentAugment hot path (hand-written sm_100a CUDA behind a
C ABI).  The reference-shaped API lives in ``latentaugment_b200.augments`` / ``.options``."""
from ._lib import LatentAugmentError  # noqa: F401

__version__ = '0.1.0'
