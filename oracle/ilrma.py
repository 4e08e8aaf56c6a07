"""Oracle for the ILRMA family (test infrastructure, see oracle/__init__.py).

Restates src/bss/ilrma.py: GaussILRMA (:178-677) and tILRMA (:713-1020).
State is a plain dict: X, W (None for ISS), Y, T, V, Z (partitioned only), pair.
"""
import numpy as np

from .core import (EPS, THRESHOLD, demix, estimate_demix_filter, projection_back_scale,
                   next_update_pair, weighted_covariance, ip_rows, ip2_pair, iss_sweep,
                   logabsdet_sum)

IP_NAMES = ('IP', 'IP1')
IP2_NAMES = ('pairwise', 'IP2')


def init_state(X, n_basis, spatial='IP', partitioning=False, W=None, T=None, V=None, Z=None,
               eps=EPS):
    """`_reset`: src/bss/ilrma.py:50-104.  Random draws use the global legacy RNG in the
    reference's order (latent, basis, activation) so np.random.seed reproduces it."""
    C, F, Tn = X.shape
    N = C
    if W is None:
        W = np.tile(np.eye(N, C, dtype=np.complex128), (F, 1, 1))
    else:
        W = W.copy()
    st = {'X': X, 'W': None if spatial == 'ISS' else W, 'Y': demix(X, W), 'pair': None}
    if partitioning:
        if Z is None:
            Z = np.random.rand(N, n_basis) * 1e-2 + 1 / N
            s = Z.sum(axis=0)
            s[s < eps] = eps
            Z = Z / s
        else:
            Z = Z.copy()
        T = np.random.rand(F, n_basis) if T is None else T.copy()
        V = np.random.rand(n_basis, Tn) if V is None else V.copy()
        st['Z'] = Z
    else:
        T = np.random.rand(N, F, n_basis) if T is None else T.copy()
        V = np.random.rand(N, n_basis, Tn) if V is None else V.copy()
    st['T'], st['V'] = T, V
    return st


def _current_estimate(st):
    if st['W'] is None:
        return st['Y']
    return demix(st['X'], st['W'])


def _model_variance(st, domain, partitioning):
    """R = (TV)^(2/domain) (:497-499) or sum_k Z T V (:493-495); not floored."""
    if partitioning:
        Z, T, V = st['Z'], st['T'], st['V']
        return np.sum(Z[:, np.newaxis, :, np.newaxis] * T[:, :, np.newaxis] * V[np.newaxis, :, :], axis=2)
    return (st['T'] @ st['V']) ** (2 / domain)


def _mu_pair(P, T, V, p, q, eps):
    """IS-type multiplicative update of (T, V) for stacked sources.  :413-428."""
    TV = T @ V
    TV[TV < eps] = eps
    num = (P / TV ** p) @ V.swapaxes(-2, -1)
    den = (1 / TV) @ V.swapaxes(-2, -1)
    den[den < eps] = eps
    T = T * (num / den) ** q
    TV = T @ V
    TV[TV < eps] = eps
    num = T.swapaxes(-2, -1) @ (P / TV ** p)
    den = T.swapaxes(-2, -1) @ (1 / TV)
    den[den < eps] = eps
    V = V * (num / den) ** q
    return T, V


def source_model_basic(st, domain=2, partitioning=False, eps=EPS):
    """src/bss/ilrma.py:356-430."""
    P = np.abs(_current_estimate(st)) ** 2
    if not partitioning:
        st['T'], st['V'] = _mu_pair(P, st['T'], st['V'], (domain + 2) / domain, domain / (domain + 2), eps)
        return
    assert domain == 2, "Not support domain = {}".format(domain)
    Z, T, V = st['Z'], st['T'], st['V']

    def sigma(Z, T, V):
        S = (Z[:, np.newaxis, :] * T[np.newaxis]) @ V[np.newaxis]
        S[S < eps] = eps
        return S

    # latent (:374-383) -- assigned, not multiplied into the old Z
    S = sigma(Z, T, V)
    TVk = T[:, :, np.newaxis] * V[np.newaxis, :, :]                        # (F,K,T)
    num = np.sum((P / S ** 2)[:, :, np.newaxis, :] * TVk, axis=(1, 3))
    den = np.sum((1 / S)[:, :, np.newaxis, :] * TVk, axis=(1, 3))
    den[den < eps] = eps
    Z = np.sqrt(num / den)
    Z = Z / Z.sum(axis=0)
    # basis (:386-394)
    S = sigma(Z, T, V)
    ZV = Z[:, :, np.newaxis] * V[np.newaxis]                               # (N,K,T)
    num = np.sum((P / S ** 2)[:, :, np.newaxis, :] * ZV[:, np.newaxis], axis=(0, 3))
    den = np.sum((1 / S)[:, :, np.newaxis, :] * ZV[:, np.newaxis], axis=(0, 3))
    den[den < eps] = eps
    T = T * np.sqrt(num / den)
    # activation (:397-405)
    S = sigma(Z, T, V)
    ZT = Z[:, np.newaxis, :] * T[np.newaxis]                               # (N,F,K)
    num = np.sum((P / S ** 2)[:, :, np.newaxis, :] * ZT[..., np.newaxis], axis=(0, 1))
    den = np.sum((1 / S)[:, :, np.newaxis, :] * ZT[..., np.newaxis], axis=(0, 1))
    den[den < eps] = eps
    V = V * np.sqrt(num / den)
    st['Z'], st['T'], st['V'] = Z, T, V


def source_model_pairwise(st, domain=2, eps=EPS):
    """Only the two sources of the current pair are updated.  src/bss/ilrma.py:432-481."""
    m, n = st['pair']
    Y = _current_estimate(st)
    T, V = st['T'], st['V']
    for s in (m, n):   # the two updates are independent of each other
        T[s], V[s] = _mu_pair(np.abs(Y[s]) ** 2, T[s], V[s], (domain + 2) / domain, domain / (domain + 2), eps)


def spatial_model_ip(st, domain=2, partitioning=False, eps=EPS, threshold=THRESHOLD):
    """src/bss/ilrma.py:483-535."""
    R = _model_variance(st, domain, partitioning)
    R[R < eps] = eps
    U = weighted_covariance(st['X'], R)
    gate = ip_rows(st['W'], U, threshold)
    st['ip_gate'] = gate   # (N,F) condition-gate decisions of this update, for the index-parity tests
    st['Y'] = demix(st['X'], st['W'])
    return U, gate


def spatial_model_iss(st, domain=2, partitioning=False, eps=EPS):
    """src/bss/ilrma.py:537-564."""
    R = _model_variance(st, domain, partitioning)
    R[R < eps] = eps
    st['Y'] = iss_sweep(st['Y'], R)


def spatial_model_pairwise(st, domain=2, partitioning=False, eps=EPS, threshold=THRESHOLD):
    """src/bss/ilrma.py:566-633."""
    R = _model_variance(st, domain, partitioning)
    m, n = st['pair']
    R_m, R_n = R[m], R[n]
    R_m[R_m < eps] = eps
    R_n[R_n < eps] = eps
    U_m = weighted_covariance(st['X'], R_m)
    U_n = weighted_covariance(st['X'], R_n)
    info = ip2_pair(st['W'], U_m, U_n, m, n, threshold)
    st['ip2_info'] = info   # (order, gate_m, gate_n, eigenvalues) of this update, for the index-parity tests
    st['Y'] = demix(st['X'], st['W'])
    return info


def normalize(st, mode='power', domain=2, partitioning=False, reference_id=0, eps=EPS):
    """src/bss/ilrma.py:293-338."""
    X = st['X']
    if st['W'] is None:
        Y = st['Y']
        W = estimate_demix_filter(Y, X)
    else:
        W = st['W']
        Y = demix(X, W)
    T = st['T']
    if mode == 'power':
        aux = np.sqrt((np.abs(Y) ** 2).mean(axis=(1, 2)))
        aux[aux < eps] = eps
        W = W / aux[np.newaxis, :, np.newaxis]
        Y = Y / aux[:, np.newaxis, np.newaxis]
        if partitioning:
            Zaux = st['Z'] / (aux[:, np.newaxis] ** domain)
            Zs = np.sum(Zaux, axis=0)
            T = T * Zs
            st['Z'] = Zaux / Zs
        else:
            T = T / (aux[:, np.newaxis, np.newaxis] ** domain)
    elif mode == 'projection-back':
        if partitioning:
            raise NotImplementedError("Not support 'projection-back' based normalization for partitioninig function. Choose 'power' based normalization.")
        scale = projection_back_scale(Y, X[reference_id])
        Y = Y * scale[..., np.newaxis]
        W = W * scale.transpose(1, 0)[..., np.newaxis]
        T = T * np.abs(scale[..., np.newaxis]) ** domain
    else:
        raise ValueError("Not support normalization based on {}. Choose 'power' or 'projection-back'".format(mode))
    st['Y'], st['T'] = Y, T
    if st['W'] is not None:
        st['W'] = W


def update_once(st, spatial='IP', domain=2, normalize_mode='power', partitioning=False,
                reference_id=0, eps=EPS, threshold=THRESHOLD):
    """GaussILRMA.update_once: src/bss/ilrma.py:286-338."""
    if spatial in IP2_NAMES:
        source_model_pairwise(st, domain, eps)
    else:
        source_model_basic(st, domain, partitioning, eps)
    if spatial in IP_NAMES:
        spatial_model_ip(st, domain, partitioning, eps, threshold)
    elif spatial == 'ISS':
        spatial_model_iss(st, domain, partitioning, eps)
    elif spatial in IP2_NAMES:
        spatial_model_pairwise(st, domain, partitioning, eps, threshold)
    else:
        raise NotImplementedError("Not support {}-based spatial update.".format(spatial))
    if normalize_mode:
        normalize(st, normalize_mode, domain, partitioning, reference_id, eps)


def negative_loglikelihood(st, domain=2, partitioning=False, eps=EPS):
    """src/bss/ilrma.py:648-677."""
    X = st['X']
    if st['W'] is None:
        Y = st['Y']
        W = estimate_demix_filter(Y, X)
    else:
        W = st['W']
        Y = demix(X, W)
    P = np.abs(Y) ** 2
    R = _model_variance(st, domain, partitioning)
    R[R < eps] = eps
    return np.sum(P / R + np.log(R)) - 2 * X.shape[2] * logabsdet_sum(W)


def run(X, iteration=100, n_basis=10, spatial='IP', domain=2, normalize_mode='power',
        partitioning=False, reference_id=0, eps=EPS, threshold=THRESHOLD,
        record_loss=True, on_iteration=None, **presets):
    """GaussILRMA.__call__: src/bss/ilrma.py:203-273.  Returns (output, state, loss list)."""
    st = init_state(X, n_basis, spatial, partitioning, eps=eps, **presets)
    loss = [negative_loglikelihood(st, domain, partitioning, eps)] if record_loss else None
    for _ in range(iteration):
        if spatial in IP2_NAMES:
            st['pair'] = next_update_pair(st['pair'], X.shape[0])
        update_once(st, spatial, domain, normalize_mode, partitioning, reference_id, eps, threshold)
        if record_loss:
            loss.append(negative_loglikelihood(st, domain, partitioning, eps))
        if on_iteration is not None:
            on_iteration(st)
    if spatial == 'ISS':
        Y = st['Y']
        st['W_final'] = estimate_demix_filter(Y, X)
    else:
        Y = demix(X, st['W'])
    out = Y * projection_back_scale(Y, X[reference_id])[..., np.newaxis]
    st['Y'] = out
    return out, st, loss


# --------------------------------------------------------------------------- t-ILRMA

def t_source_model(st, nu, eps=EPS):
    """src/bss/ilrma.py:859-938 (non-partitioned branch :915-938, domain = 2 only)."""
    P = np.abs(_current_estimate(st)) ** 2
    T, V = st['T'], st['V']
    c = 2 + nu

    def stats(T, V):
        TV = T @ V
        TV[TV < eps] = eps
        with np.errstate(divide='ignore'):
            h = 1 / (2 / (c * TV) + nu / (c * P))
        return h / TV ** 2, 1 / TV

    d, inv = stats(T, V)
    den = inv @ V.swapaxes(-2, -1)
    den[den < eps] = eps
    T = T * np.sqrt(d @ V.swapaxes(-2, -1) / den)
    d, inv = stats(T, V)
    den = T.swapaxes(-2, -1) @ inv
    den[den < eps] = eps
    V = V * np.sqrt(T.swapaxes(-2, -1) @ d / den)
    st['T'], st['V'] = T, V


def t_spatial_model(st, nu, eps=EPS):
    """src/bss/ilrma.py:940-991: weight xi = (nu R + 2 P)/(nu + 2), plain inverse (no
    condition gate) and a floored denominator (:977-982)."""
    X, W = st['X'], st['W']
    P = np.abs(demix(X, W)) ** 2
    R = st['T'] @ st['V']
    R[R < eps] = eps
    U = weighted_covariance(X, (nu * R + 2 * P) / (nu + 2))
    for n in range(W.shape[1]):
        w = np.linalg.inv(W @ U[n])[..., n]
        q = w[:, np.newaxis, :].conj() @ U[n] @ w[:, :, np.newaxis]
        den = np.sqrt(q.squeeze(axis=-1))
        den[den < eps] = eps
        W[:, n, :] = w.conj() / den
    st['Y'] = demix(X, W)


def t_update_once(st, nu, normalize_mode='power', eps=EPS):
    """tILRMA.update_once: src/bss/ilrma.py:814-857 (power normalisation, exponent 2)."""
    t_source_model(st, nu, eps)
    t_spatial_model(st, nu, eps)
    if normalize_mode:
        if normalize_mode != 'power':
            raise ValueError("Not support normalization based on {}. Choose 'power' or 'projection-back'".format(normalize_mode))
        Y = demix(st['X'], st['W'])
        aux = np.sqrt((np.abs(Y) ** 2).mean(axis=(1, 2)))
        aux[aux < eps] = eps
        st['W'] = st['W'] / aux[np.newaxis, :, np.newaxis]
        st['Y'] = Y / aux[:, np.newaxis, np.newaxis]
        st['T'] = st['T'] / (aux[:, np.newaxis, np.newaxis] ** 2)


def t_negative_loglikelihood(st, nu, eps=EPS):
    """src/bss/ilrma.py:993-1020."""
    X, W = st['X'], st['W']
    P = np.abs(demix(X, W)) ** 2
    R = st['T'] @ st['V']
    R[R < eps] = eps
    return (np.sum((1 + nu / 2) * np.log(1 + (2 / nu) * (P / R)) + np.log(R))
            - 2 * X.shape[2] * logabsdet_sum(W))


def t_run(X, iteration=100, n_basis=10, nu=1, normalize_mode='power', reference_id=0,
          eps=EPS, record_loss=True, **presets):
    """tILRMA.__call__: src/bss/ilrma.py:733-800."""
    st = init_state(X, n_basis, 'IP', False, eps=eps, **presets)
    loss = [t_negative_loglikelihood(st, nu, eps)] if record_loss else None
    for _ in range(iteration):
        t_update_once(st, nu, normalize_mode, eps)
        if record_loss:
            loss.append(t_negative_loglikelihood(st, nu, eps))
    Y = demix(X, st['W'])
    out = Y * projection_back_scale(Y, X[reference_id])[..., np.newaxis]
    st['Y'] = out
    return out, st, loss
