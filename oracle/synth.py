"""Seeded synthetic inputs shared by tests, smoke() and bench.py (test infrastructure).

`mix2` is the generator of SURVEY.md Appendix D; it produces a well-conditioned determined
mixture of low-rank-variance Gaussian sources.  All generators return complex128 values that are
exactly representable in complex64, so the CUDA path (complex64 storage) and the oracle
(complex128) see identical numbers.
"""
import numpy as np


def _c64_exact(X):
    return X.astype(np.complex64).astype(np.complex128)


def mix2(C, F, T, K=2, seed=0, snr_db=30.0, lo=0.05):
    rng = np.random.default_rng(seed)
    Tb = lo + rng.random((C, F, K))
    Vb = lo + rng.random((C, K, T)) ** 2
    R = Tb @ Vb
    S = np.sqrt(R / 2) * (rng.standard_normal((C, F, T)) + 1j * rng.standard_normal((C, F, T)))
    A = np.eye(C) + 0.5 * (rng.standard_normal((F, C, C)) + 1j * rng.standard_normal((F, C, C))) / np.sqrt(2)
    X = (A @ S.transpose(1, 0, 2)).transpose(1, 0, 2)
    p = np.mean(np.abs(X) ** 2)
    Nz = np.sqrt(p * 10 ** (-snr_db / 10) / 2) * (rng.standard_normal((C, F, T)) + 1j * rng.standard_normal((C, F, T)))
    return _c64_exact(X + Nz)


def iid(C, F, T, seed=0):
    rng = np.random.default_rng(seed)
    return _c64_exact((rng.standard_normal((C, F, T)) + 1j * rng.standard_normal((C, F, T))) / np.sqrt(2))


def initial_state(C, F, T, K, seed=7, n_sources=None, round32=True):
    """W = I, T0 ~ U(0,1) (N,F,K), V0 ~ U(0,1) (N,K,T) drawn in that order.
    With round32 the values are rounded to float32-representable numbers (what the CUDA path
    stores); round32=False reproduces the Appendix D known-answer runs exactly."""
    N = C if n_sources is None else n_sources
    rng = np.random.default_rng(seed)
    W = np.tile(np.eye(C, dtype=np.complex128), (F, 1, 1))
    T0 = rng.random((N, F, K))
    V0 = rng.random((N, K, T))
    if round32:
        T0 = T0.astype(np.float32).astype(np.float64)
        V0 = V0.astype(np.float32).astype(np.float64)
    return W, T0, V0


def random_demix(C, F, seed=3, spread=0.3):
    """Well-conditioned random demixing filters (F,C,C), complex64-representable."""
    rng = np.random.default_rng(seed)
    W = np.eye(C) + spread * (rng.standard_normal((F, C, C)) + 1j * rng.standard_normal((F, C, C)))
    return _c64_exact(W)


def spectrogram(F, T, seed=0):
    """Non-negative target for the NMF family: |CN(0,1)|^2, float32-representable."""
    rng = np.random.default_rng(seed)
    Z = (rng.standard_normal((F, T)) ** 2 + rng.standard_normal((F, T)) ** 2) / 2
    return Z.astype(np.float32).astype(np.float64)


def toy_dnn(gain=0.8, bias=1e-3, width=5):
    """A deterministic stand-in for the source-model DNN of GaussIDLMA (src/sss/idlma.py:212-226): a torch module with one
    parameter that smooths each (source, bin) row over `width` frames.  Tests and the pinning script build it identically."""
    import torch

    class ToyDNN(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.gain = torch.nn.Parameter(torch.tensor([float(gain)]))

        def forward(self, x):
            y = torch.nn.functional.avg_pool1d(x, kernel_size=width, stride=1, padding=width // 2, count_include_pad=False)
            return self.gain * y + bias

    return ToyDNN()


def dnn_as_callable(module):
    """What the reference's estimate_by_dnn does around the module (src/sss/idlma.py:216-224): float32 in, float32 out."""
    import torch

    def call(a):
        with torch.no_grad():
            return module(torch.Tensor(a)).cpu().numpy()
    return call
