"""diffmm_b200 — B200-native (sm_100a) implementation of the DiffMM data-parallel hot path.

Layout
  csrc/            hand-written CUDA kernels + the C ABI (include/diffmm_b200.h)
  _lib.py, ops.py  ctypes binding and tensor-level wrappers
  Model.py ...     host-side mirror of the reference call surface (Model / Denoise /
                   GaussianDiffusion / Utils.Utils / DataHandler / Conf / Main), see dropin/
There is no CPU fallback: importing is cheap, but every operator raises if libdiffmm_b200.so
is missing or the tensors are not on a CUDA device.
"""
__version__ = "0.1.0"
