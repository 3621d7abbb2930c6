"""nhvr_b200 — B200 (sm_100a) native rendering hot path of Neural-Human-Video-Rendering.

Host side is Python/PyTorch (device memory, streams, torch.distributed); all compute on the path runs
in hand-written CUDA behind the C-ABI in include/nhvr.h (libnhvr_sm100.so).  There is no CPU or
PyTorch fallback: importing ``nhvr_b200.capi`` without the built library, or calling an op off an
sm_100 device, raises.
"""
__version__ = "0.1.0"
