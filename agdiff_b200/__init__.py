"""agdiff_b200 -- B200-native conformer-sampling hot path of AGDIFF.

Host-side mirror of the reference interface (``epsnet.get_model``, ``DualEncoderEpsNetwork``)
over hand-written sm_100a CUDA kernels behind the C ABI in ``include/agdiff_b200.h``.
Importing the package does not load the native library; the first compute call does, and fails
loudly if it is missing (there is no CPU or PyTorch fallback).
"""
from .epsnet import DualEncoderEpsNetwork, get_model  # noqa: F401

__all__ = ["get_model", "DualEncoderEpsNetwork"]
__version__ = "0.1.0"
