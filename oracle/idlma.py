"""Oracle for Gauss-IDLMA (test infrastructure, see oracle/__init__.py).

Restates src/sss/idlma.py: GaussIDLMA (:88-258).  The DNN is any callable `dnn(a)` mapping a float64 array
(n_sources, n_bins, n_frames) to an array of the same shape (the reference wraps a torch module: float32 in, float32
out, src/sss/idlma.py:216-224).  State dict: X, W, dnn_output.
"""
import numpy as np

from .core import (EPS, THRESHOLD, demix, estimate_demix_filter, projection_back_scale, weighted_covariance, ip_rows,
                   logabsdet_sum)


def init_state(X):
    """src/sss/idlma.py:20-39: W = I for every bin, dnn_output = 1 (presets are not honoured)."""
    C, F, T = X.shape
    return {'X': X, 'W': np.tile(np.eye(C, C, dtype=np.complex128), (F, 1, 1)), 'dnn_output': np.ones((C, F, T))}


def update_source_model(st, dnn, domain=2, dnn_flooring=1e-5):
    """src/sss/idlma.py:167-173, :212-232."""
    P = np.abs(demix(st['X'], st['W'])) ** 2
    out = np.asarray(dnn(P ** (domain / 2))) ** (2 / domain)
    if dnn_flooring:
        out = np.maximum(out, dnn_flooring)
    st['dnn_output'] = out


def variance(st, domain=2, eps=EPS):
    """R = dnn_output^(2/domain) in the precision of dnn_output, floored at eps (src/sss/idlma.py:181,190 and :253-254)."""
    R = st['dnn_output'] ** (2 / domain)
    R[R < eps] = eps
    return R


def update_space_model(st, domain=2, eps=EPS, threshold=THRESHOLD):
    """src/sss/idlma.py:175-210."""
    U = weighted_covariance(st['X'], variance(st, domain, eps))
    return U, ip_rows(st['W'], U, threshold)


def normalize_projection_back(st, reference_id=0):
    """src/sss/idlma.py:150-158: scale the estimates, then re-derive W by least squares."""
    X = st['X']
    Y = demix(X, st['W'])
    Y = Y * projection_back_scale(Y, X[reference_id])[..., np.newaxis]
    st['W'] = estimate_demix_filter(Y, X)
    return Y


def update_once(st, dnn, domain=2, reference_id=0, dnn_flooring=1e-5, eps=EPS, threshold=THRESHOLD,
                is_source_model_update=True):
    """src/sss/idlma.py:142-165 with normalize='projection-back' (the only branch that does not raise)."""
    if is_source_model_update:
        update_source_model(st, dnn, domain, dnn_flooring)
    update_space_model(st, domain, eps, threshold)
    return normalize_projection_back(st, reference_id)


def negative_loglikelihood(st, domain=2, eps=EPS):
    """src/sss/idlma.py:244-258."""
    X, W = st['X'], st['W']
    P = np.abs(demix(X, W)) ** 2
    R = variance(st, domain, eps)
    return np.sum(P / R + np.log(R)) - 2 * X.shape[2] * logabsdet_sum(W)


def run(X, dnn, iteration=100, domain=2, reference_id=0, dnn_flooring=1e-5, eps=EPS, threshold=THRESHOLD):
    """GaussIDLMA.__call__ (src/sss/idlma.py:105-140)."""
    st = init_state(X)
    loss = [negative_loglikelihood(st, domain, eps)]
    for _ in range(iteration):
        update_once(st, dnn, domain, reference_id, dnn_flooring, eps, threshold)
        loss.append(negative_loglikelihood(st, domain, eps))
    Y = demix(X, st['W'])
    out = Y * projection_back_scale(Y, X[reference_id])[..., np.newaxis]
    return out, st, loss
