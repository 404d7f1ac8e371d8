"""glam_b200 — B200-native (sm_100a) implementation of GLAM's molecular message-passing hot path.

Host side: drop-in mirrors of the reference's `layer.py` names (`glam_b200.layer`) and model wiring
(`glam_b200.model`); device side: hand-written CUDA behind the C ABI in include/glam_b200.h.
"""
ABI_VERSION = 14
__version__ = "0.1.0"
