"""NumPy-1.x semantics for np.linalg.solve(A (...,C,C), b (...,C)) -- pinning tool only.

The reference was written for NumPy 1.x, where a `b` with one dimension less than `A` is a
stack of vectors (src/bss/ilrma.py:523, src/bss/iva.py:511,744, src/bss/mnmf.py:880).  NumPy >= 2
rejects that call, so the unmodified reference cannot run without this shim.  Import it before
importing the reference; the reference files themselves are never modified or copied.
"""
import numpy as np

_orig_solve = np.linalg.solve


def _solve(a, b):
    a = np.asarray(a)
    b = np.asarray(b)
    if b.ndim == a.ndim - 1:
        return _orig_solve(a, b[..., None])[..., 0]
    return _orig_solve(a, b)


np.linalg.solve = _solve
