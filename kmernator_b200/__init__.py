"""kmernator_b200 -- B200-native k-mer spectrum hot path (count table, histogram, FilterReads lookup pass).

Layout:
  csrc/                 hand-written sm_100a CUDA kernels + the C ABI of include/kmernator_b200.h
  capi.py               ctypes binding of the C ABI (used by tests and bench.py)
  host/                 C++ host mirror of the reference's KmerSpectrum / ReadSet / ReadSelector / FilterReads surface

There is no CPU fallback: importing `kmernator_b200.capi.load()` raises if the CUDA library is not built.
"""
from . import capi  # noqa: F401
from .capi import Context, KmnError  # noqa: F401
