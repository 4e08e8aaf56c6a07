"""B200-native update loop for STFT-domain blind source separation.

Drop-in classes for the iterative paths of tky823/audio_source_separation
(`bss.ilrma`, `bss.iva`, `bss.mnmf`, `algorithm.nmf`): same constructors, `__call__`, `update_once`,
`separate`, losses and public state attributes as the reference, every update executed by
hand-written sm_100a CUDA kernels behind the C ABI of `libbssgpu.so` (include/bssgpu.h).
There is no CPU fallback.
"""
from . import _lib  # noqa: F401

__version__ = '0.1.0'
