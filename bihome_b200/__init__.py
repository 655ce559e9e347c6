"""bihome_b200 -- B200-native implementation of the biHomE training hot path.

The compute lives in ``libbihome_b200.so`` (hand-written sm_100a CUDA behind the C ABI declared in
``include/bihome_b200.h``); this package is the PyTorch-facing host side that mirrors the reference's
module/function signatures (``src/data/utils.py``, ``src/heads/PerceptualHead.py``,
``src/heads/ransac_utils.py``).  There is no CPU fallback: importing ``bihome_b200.cabi`` without the
built library, or calling an op on a non-CUDA tensor, raises.
"""
__version__ = '0.1.0'
