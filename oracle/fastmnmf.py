"""Oracle for FastMNMF (test infrastructure, see oracle/__init__.py).

Restates src/bss/mnmf.py: FastMultichannelISNMF (:637-946), non-partitioned path.
State dict: X (M,F,T), Q (F,M,M) diagonaliser, G (N,F,M) diagonal spatial
covariances, W (N,F,K) basis, H (N,K,T) activation.
"""
import numpy as np

from .core import EPS, THRESHOLD, outer_products, solve_vec


def init_state(X, n_basis, n_sources=None, W=None, H=None):
    """src/bss/mnmf.py:653-689.  Q and G are always re-created (:660-663)."""
    M, F, Tn = X.shape
    N = M if n_sources is None else n_sources
    Q = np.tile(np.eye(M, dtype=np.complex128), (F, 1, 1))
    G = np.ones((N, F, M)) * 1e-2
    for m in range(M):
        G[m % N, :, m] = 1
    W = np.random.rand(N, F, n_basis) if W is None else W.copy()
    H = np.random.rand(N, n_basis, Tn) if H is None else H.copy()
    return {'X': X, 'Q': Q, 'G': G, 'W': W, 'H': H}


def diagonalised_power(X, Q):
    """x~[f,t,m] = |sum_c Q[f,m,c] X[c,f,t]|^2.  src/bss/mnmf.py:782-783."""
    Xf = X.transpose(1, 2, 0)
    QX = np.sum(Q[:, np.newaxis, :, :] * Xf[:, :, np.newaxis, :], axis=3)
    return np.abs(QX) ** 2


def _variance(W, H, G, eps):
    """Lambda = W H; R[f,t,m] = max(sum_n Lambda g, eps).  :790-792."""
    Lam = W @ H
    R = np.sum(Lam[..., np.newaxis] * G[:, :, np.newaxis], axis=0)
    R[R < eps] = eps
    return Lam, R


def update_nmf(st, eps=EPS):
    """src/bss/mnmf.py:775-815."""
    G, W, H = st['G'], st['W'], st['H']
    xt = diagonalised_power(st['X'], st['Q'])

    def ratios(W, H):
        _, R = _variance(W, H, G, eps)
        a = np.sum(G[:, :, np.newaxis] * (xt / R ** 2)[np.newaxis], axis=3)   # (N,F,T)
        b = np.sum(G[:, :, np.newaxis] / R[np.newaxis], axis=3)
        return a, b

    a, b = ratios(W, H)
    num = np.sum(H[:, np.newaxis, :, :] * a[:, :, np.newaxis], axis=3)
    den = np.sum(H[:, np.newaxis, :, :] * b[:, :, np.newaxis], axis=3)
    den[den < eps] = eps
    W = W * np.sqrt(num / den)
    a, b = ratios(W, H)
    num = np.sum(W[:, :, :, np.newaxis] * a[:, :, np.newaxis], axis=1)
    den = np.sum(W[:, :, :, np.newaxis] * b[:, :, np.newaxis], axis=1)
    den[den < eps] = eps
    H = H * np.sqrt(num / den)
    st['W'], st['H'] = W, H


def update_scm(st, eps=EPS):
    """src/bss/mnmf.py:817-846."""
    Lam, R = _variance(st['W'], st['H'], st['G'], eps)
    xt = diagonalised_power(st['X'], st['Q'])
    A = np.sum(Lam[..., np.newaxis] * (xt / R ** 2)[np.newaxis], axis=2)
    B = np.sum(Lam[..., np.newaxis] / R[np.newaxis], axis=2)
    B[B < eps] = eps
    st['G'] = st['G'] * np.sqrt(A / B)


def update_diagonalizer(st, eps=EPS, threshold=THRESHOLD):
    """src/bss/mnmf.py:848-888.  R is fixed during the loop over channels.
    Returns (V (M,F,M,M), gate (M,F))."""
    Q = st['Q']
    F, M = Q.shape[0], Q.shape[1]
    XX = outer_products(st['X'])
    _, R = _variance(st['W'], st['H'], st['G'], eps)
    E = np.tile(np.eye(M), (F, 1, 1))
    Vs = np.zeros((M, F, M, M), dtype=np.complex128)
    gate = np.zeros((M, F), dtype=bool)
    for m in range(M):
        V = (XX / R[:, :, m, np.newaxis, np.newaxis]).mean(axis=1)
        QV = Q @ V
        ok = np.linalg.cond(QV) < threshold
        q = solve_vec(QV, E[:, m, :])
        s = q.conj()[:, np.newaxis, :] @ V @ q[:, :, np.newaxis]
        den = np.sqrt(s[..., 0])
        den[den < eps] = eps
        Q[:, m, :] = np.where(ok[:, np.newaxis], q.conj() / den, Q[:, m, :])
        Vs[m], gate[m] = V, ok
    return Vs, gate


def normalize(st, eps=EPS):
    """src/bss/mnmf.py:745-771 (all in place in the reference)."""
    Q, G, W, H = st['Q'], st['G'], st['W'], st['H']
    s = np.real((Q * Q.conj()).sum(axis=2).mean(axis=1))
    s[s < eps] = eps
    Q /= np.sqrt(s)[:, np.newaxis, np.newaxis]
    G /= s[np.newaxis, :, np.newaxis]
    gs = G.sum(axis=2)
    gs[gs < eps] = eps
    G /= gs[:, :, np.newaxis]
    W *= gs[:, :, np.newaxis]
    ws = W.sum(axis=1)
    ws[ws < eps] = eps
    W /= ws[:, np.newaxis]
    H *= ws[:, :, np.newaxis]


def update_once(st, normalize_mode='power', eps=EPS, threshold=THRESHOLD):
    """src/bss/mnmf.py:737-773."""
    update_nmf(st, eps)
    update_scm(st, eps)
    update_diagonalizer(st, eps, threshold)
    if normalize_mode:
        if normalize_mode != 'power':
            raise ValueError("Not support normalization based on {}. Choose 'power'".format(normalize_mode))
        normalize(st, eps)


def negative_loglikelihood(st, eps=EPS):
    """src/bss/mnmf.py:890-917 (note the plain transpose in det(Q Q^T), :911)."""
    Q = st['Q']
    Lam = st['W'] @ st['H']
    yt = np.sum(Lam[..., np.newaxis] * st['G'][:, :, np.newaxis, :], axis=0) + eps
    xt = diagonalised_power(st['X'], Q) + eps
    det = np.abs(np.linalg.det(Q @ Q.transpose(0, 2, 1)))
    return np.sum(xt / yt + np.log(yt)) - st['X'].shape[2] * np.sum(np.log(det))


def separate(st, reference_id=0, eps=EPS):
    """Multichannel Wiener filter, reference-microphone image (N,F,T).  src/bss/mnmf.py:919-946."""
    X, Q = st['X'].transpose(1, 2, 0), st['Q']
    Lam = st['W'] @ st['H']
    LG = Lam[..., np.newaxis] * st['G'][:, :, np.newaxis, :]                 # (N,F,T,M)
    yt = np.sum(LG, axis=0)
    Qinv = np.linalg.inv(Q)
    QX = np.sum(Q[:, np.newaxis, :] * X[:, :, np.newaxis], axis=3)
    yt[yt < eps] = eps
    Z = QX * (LG / yt)
    xhat = np.sum(Qinv[:, np.newaxis, :, :] * Z[:, :, :, np.newaxis, :], axis=4)
    return xhat.transpose(0, 3, 1, 2)[:, reference_id, :, :]


def run(X, iteration=100, n_basis=10, n_sources=None, normalize_mode='power', reference_id=0,
        eps=EPS, threshold=THRESHOLD, record_loss=True, **presets):
    """src/bss/mnmf.py:691-722."""
    st = init_state(X, n_basis, n_sources, **presets)
    loss = [negative_loglikelihood(st, eps)] if record_loss else None
    for _ in range(iteration):
        update_once(st, normalize_mode, eps, threshold)
        if record_loss:
            loss.append(negative_loglikelihood(st, eps))
    out = separate(st, reference_id, eps)
    return out, st, loss
