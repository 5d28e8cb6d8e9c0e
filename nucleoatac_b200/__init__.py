"""nucleoatac_b200 -- B200-native per-chunk occ/nuc scoring path of NucleoATAC.

Python 3 host code over hand-written sm_100a CUDA kernels, called through the ctypes C-ABI of
``libnucleo_b200.so`` (include/nucleo_b200.h).  No torch, no Triton, no CPU fallback: importing
the compute classes needs the built library, creating an Engine needs a B200.
"""
__version__ = "0.1.0"
