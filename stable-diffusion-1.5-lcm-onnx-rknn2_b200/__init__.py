"""dreamlab_b200 — B200-native LCM denoise + VAE decode hot path (SD1.5-LCM).

Host side is Python/PyTorch (device memory, streams, CUDA graphs); all arithmetic runs in the
hand-written sm_100a kernels of `csrc/` behind the C-ABI of `include/dreamlab_b200.h`.
There is no CPU fallback: importing is cheap, but every op raises if the native library or a
CUDA device is missing.
"""
__all__ = ["lib", "scheduler", "weights", "engine", "synthetic"]
