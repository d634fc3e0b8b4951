"""dreamer_b200 -- B200-native (sm_100a) hot path of the youngers2006/Dreamer RSSM.

Package layout
  csrc/                 hand-written CUDA (TMA + tcgen05 fused GEMM stages, HBM-bound helpers) and
                        the C-ABI of include/dreamer_b200.h  ->  libdreamer_b200.so
  _lib.py, ops.py       ctypes binding and tensor-level wrappers (no CPU fallback)
  SequenceModel.py ...  mirrors of the reference modules (same class names, constructor
                        signatures and state_dict keys) whose forward passes call the library
"""
from . import _lib  # noqa: F401

__all__ = ["_lib"]
__version__ = "0.1.0"
