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

    scale = reference Y^H (Y Y^H)^-1 per bin, for arbitrary `Y` and `reference` exactly as upstream: both arrays go to
    the device and one kernel forms the two small products and the solve (`bss_least_squares_map`).

    Optional shortcut (our extension, not part of the reference signature): when the estimates are Y = demix_filter @ input
    and `reference` is a channel (or all channels) of `input`, passing `input` (C,F,T) and `demix_filter` (F,N,C) lets the
    device work from the C x C mixture covariance instead of the frames.
    """
    Y = np.asarray(Y)
    reference = np.asarray(reference)
    if reference.ndim not in (2, 3):
        raise ValueError("reference.ndim is expected 2 or 3, but given {}.".format(reference.ndim))
    if input is not None and demix_filter is not None:
        input = np.asarray(input)
        n_channels = input.shape[0]
        if reference.ndim == 2:
            ref_ids = [i for i in range(n_channels) if reference is input[i] or np.array_equal(reference, input[i])]
            if ref_ids:
                return _lib.projection_back_scale(input, demix_filter, ref_ids[0])
        elif reference.shape == input.shape and (reference is input or np.array_equal(reference, input)):
            return np.stack([_lib.projection_back_scale(input, demix_filter, c) for c in range(n_channels)], axis=0)
    if reference.ndim == 2:
        return _lib.least_squares_map(reference[np.newaxis], Y)[0]
    return _lib.least_squares_map(reference, Y)
