"""Oracle for auxiliary-function IVA (test infrastructure, see oracle/__init__.py).

Restates src/bss/iva.py: AuxLaplaceIVA (:388-619) and AuxGaussIVA (:621-802).
`kind` is 'laplace' or 'gauss'.  State dict: X, W (None for ISS), Y, pair.
"""
import numpy as np

from .core import (EPS, THRESHOLD, demix, estimate_demix_filter, projection_back_scale,
                   next_update_pair, weighted_covariance, ip_rows, ip2_pair, iss_sweep,
                   logabsdet_sum)

IP_NAMES = ('IP', 'IP1')
IP2_NAMES = ('pairwise', 'IP2')


def init_state(X, spatial='IP', W=None):
    """src/bss/iva.py:39-59 and :356-360."""
    C, F, _ = X.shape
    W = np.tile(np.eye(C, C, dtype=np.complex128), (F, 1, 1)) if W is None else W.copy()
    return {'X': X, 'W': None if spatial == 'ISS' else W, 'Y': demix(X, W), 'pair': None}


def frame_weights(Y, kind):
    """r[n,t]: Laplace sqrt(sum_f |y|^2) (src/bss/iva.py:489-490), Gauss mean_f |y|^2 (:722-723).
    Not floored."""
    P = np.abs(Y) ** 2
    if kind == 'laplace':
        return np.sqrt(P.sum(axis=1))
    if kind == 'gauss':
        return P.mean(axis=1)
    raise ValueError(kind)


def update_ip(st, kind, eps=EPS, threshold=THRESHOLD):
    """src/bss/iva.py:481-523 / :714-756.  Weights come from the stored estimate."""
    r = frame_weights(st['Y'], kind)[:, np.newaxis, :]
    r[r < eps] = eps
    U = weighted_covariance(st['X'], r)
    gate = ip_rows(st['W'], U, threshold)
    st['Y'] = demix(st['X'], st['W'])
    return U, gate


def update_iss(st, kind, eps=EPS):
    """src/bss/iva.py:525-542 / :758-775."""
    r = frame_weights(st['Y'], kind)
    r[r < eps] = eps
    st['Y'] = iss_sweep(st['Y'], r[:, np.newaxis, :])


def update_pairwise(st, kind, eps=EPS, threshold=THRESHOLD):
    """src/bss/iva.py:544-599 (Laplace only; AuxGaussIVA raises NotImplementedError, :777-778)."""
    if kind != 'laplace':
        raise NotImplementedError("In progress...")
    m, n = st['pair']
    Y = st['Y']
    r_m = np.sqrt((np.abs(Y[m]) ** 2).sum(axis=0))[np.newaxis, :]
    r_n = np.sqrt((np.abs(Y[n]) ** 2).sum(axis=0))[np.newaxis, :]
    r_m[r_m < eps] = eps
    r_n[r_n < eps] = eps
    U_m = weighted_covariance(st['X'], r_m)
    U_n = weighted_covariance(st['X'], r_n)
    info = ip2_pair(st['W'], U_m, U_n, m, n, threshold)
    st['ip2_info'] = info   # (order, gate_m, gate_n, eigenvalues) of this update, for the index-parity tests
    st['Y'] = demix(st['X'], st['W'])
    return info


def update_once(st, kind, spatial='IP', eps=EPS, threshold=THRESHOLD):
    """src/bss/iva.py:469-479 / :702-712."""
    if spatial in IP_NAMES:
        update_ip(st, kind, eps, threshold)
    elif spatial == 'ISS':
        update_iss(st, kind, eps)
    elif spatial in IP2_NAMES:
        update_pairwise(st, kind, eps, threshold)
    else:
        raise ValueError("Not support {} based spatial updates.".format(spatial))


def negative_loglikelihood(st, kind, eps=EPS):
    """Laplace: src/bss/iva.py:604-619.  Gauss: :783-802."""
    X = st['X']
    if st['W'] is None:
        Y = st['Y']
        W = estimate_demix_filter(Y, X)
    else:
        W = st['W']
        Y = demix(X, W)
    n_bins, n_frames = X.shape[1], X.shape[2]
    if kind == 'laplace':
        return (2 * np.sqrt(np.sum(np.abs(Y) ** 2, axis=1))).sum() - 2 * n_frames * logabsdet_sum(W)
    Y = demix(X, W)                                   # :796 recomputes from W
    R = (np.abs(Y) ** 2).mean(axis=1)
    R[R < eps] = eps
    return n_bins * np.sum(np.log(R)) - 2 * n_frames * logabsdet_sum(W)


def run(X, iteration=100, kind='laplace', spatial='IP', reference_id=0, apply_projection_back=True,
        eps=EPS, threshold=THRESHOLD, record_loss=True, W=None):
    """AuxLaplaceIVA.__call__ (:392-460) / AuxGaussIVA.__call__ (:625-693)."""
    st = init_state(X, spatial, W)
    loss = [negative_loglikelihood(st, kind, eps)] if record_loss else None
    for _ in range(iteration):
        if spatial in IP2_NAMES:
            st['pair'] = next_update_pair(st['pair'], X.shape[0])
        update_once(st, kind, spatial, eps, threshold)
        if record_loss:
            loss.append(negative_loglikelihood(st, kind, eps))
    if spatial == 'ISS':
        Y = st['Y']
        st['W_final'] = estimate_demix_filter(Y, X)
    else:
        Y = demix(X, st['W'])
    out = Y
    if apply_projection_back:
        out = out * projection_back_scale(out, X[reference_id])[..., np.newaxis]
    st['Y'] = out
    return out, st, loss
