"""STFT feed on the GPU with the reference's function surface (src/transform/stft.py == src/algorithm/stft.py).

`stft(input, fft_size, hop_size, window_fn)` and `istft(...)` reproduce `scipy.signal.stft / istft` as the reference
calls them (`nperseg=fft_size, noverlap=fft_size-hop_size, window=window_fn`, everything else default: zero boundary
extension, tail padding, one-sided spectrum, 'spectrum' scaling); framing, windowing, the FFT itself and the overlap-add
run in CUDA kernels (float32 arithmetic, results returned as complex128 / float64).  The window coefficients are host
plumbing (`scipy.signal.get_window`, what scipy itself uses for a window name).  `fft_size` must be a power of two.
"""
import numpy as np
from scipy import signal as ss

from .. import _lib


def _window(fft_size, window_fn):
    if isinstance(window_fn, (str, tuple)):
        return np.asarray(ss.get_window(window_fn, fft_size), dtype=np.float64)
    window = np.asarray(window_fn, dtype=np.float64)
    if window.shape != (fft_size,):
        raise ValueError('window must have length nperseg')
    return window


def stft(input, fft_size, hop_size=None, window_fn='hann', normalize=False):
    """src/transform/stft.py:4-8.  input (..., n_samples) -> (..., fft_size // 2 + 1, n_frames) complex128."""
    return _lib.stft(input, fft_size, hop_size, _window(fft_size, window_fn))


def istft(input, fft_size, hop_size=None, window_fn='hann', normalize=False, length=None):
    """src/transform/stft.py:10-17."""
    output = _lib.istft(input, fft_size, hop_size, _window(fft_size, window_fn))

    if length is not None:
        output = output[..., :length]

    return output


def build_window(fft_size, window_fn='hann'):
    """src/transform/stft.py:19-27 (periodic windows)."""
    if window_fn == 'hann':
        window = ss.get_window('hann', fft_size)
    elif window_fn == 'hamming':
        window = ss.get_window('hamming', fft_size)
    else:
        raise ValueError("Not support {} window.".format(window_fn))

    return window


def build_optimal_window(window, hop_size=None):
    """src/transform/stft.py:29-48."""
    window_length = len(window)

    if hop_size is None:
        hop_size = window_length // 2

    windows = np.concatenate([
        np.roll(window[np.newaxis, :], hop_size * idx) for idx in range(window_length // hop_size)
    ], axis=0)

    power = windows**2
    norm = power.sum(axis=0)
    optimal_window = window / norm

    return optimal_window
