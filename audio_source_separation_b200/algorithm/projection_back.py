"""projection_back on the GPU (src/algorithm/projection_back.py:3-34)."""
import numpy as np

from .. import _lib


def projection_back(Y, reference, input=None, demix_filter=None):
    """
    Args:
        Y: (n_sources, n_bins, n_frames)
        reference: (n_bins, n_frames) or (n_channels, n_bins, n_frames)
    Returns:
        scale: (n_sources, n_bins) or (n_channels, n_sources, n_bins)

    The device kernel works from the demixing filter and the mixture covariance, so the estimates are
    described by `input` (C,F,T) and `demix_filter` (F,N,C) with Y = demix_filter @ input; when they are
    not given, Y itself is used as the mixture with an identity filter.
    """
    Y = np.asarray(Y)
    reference = np.asarray(reference)
    if reference.ndim not in (2, 3):
        raise ValueError("reference.ndim is expected 2 or 3, but given {}.".format(reference.ndim))
    if input is None or demix_filter is None:
        raise NotImplementedError("projection_back on the device needs `input` and `demix_filter` (Y = W X)")
    n_channels = input.shape[0]
    if reference.ndim == 2:
        ref_ids = [i for i in range(n_channels) if reference is input[i] or np.array_equal(reference, input[i])]
        if not ref_ids:
            raise ValueError("reference must be one of the input channels")
        return _lib.projection_back_scale(input, demix_filter, ref_ids[0])
    return np.stack([_lib.projection_back_scale(input, demix_filter, c) for c in range(n_channels)], axis=0)
