"""Oracle for Sawada's multichannel IS-NMF (test infrastructure, see oracle/__init__.py).

Restates src/bss/mnmf.py: MultichannelISNMF (:116-635) with author='Sawada' (the Ozerov EM variant is marked
"in progress" upstream and is out of scope), together with the helpers it calls:
`solve_Riccati` (src/algorithm/linalg.py:7-30), `to_PSD` (src/utils/utils_linalg.py:9-31) and
`logdet_divergence` (src/criterion/divergence.py:83-106).
State dict: X (C,F,T) mixture, H (F,N,C,C) spatial covariances, Z (N,K) latent, T (F,K) basis, V (K,T) activation.
"""
import numpy as np

from .core import EPS


def init_state(X, n_basis, n_sources=None, H=None, Z=None, T=None, V=None, eps=EPS):
    """src/bss/mnmf.py:202-237: latent, spatial, basis, activation are drawn in that order."""
    C, F, Tn = X.shape
    N = C if n_sources is None else n_sources
    if Z is None:
        Z = np.random.rand(N, n_basis) * 1e-2 + 1 / N
        s = Z.sum(axis=0)
        s[s < eps] = eps
        Z = Z / s
    else:
        Z = Z.copy()
    H = np.tile(np.eye(C), (F, N, 1, 1)) if H is None else H.copy()
    T = np.random.rand(F, n_basis) if T is None else T.copy()
    V = np.random.rand(n_basis, Tn) if V is None else V.copy()
    return {'X': X, 'H': H, 'Z': Z, 'T': T, 'V': V}


def source_power(st):
    """lambda[n,f,t] = sum_k Z[n,k] T[f,k] V[k,t] (the `ZTV` of :624, the weights of :554-560)."""
    return np.einsum('nk,fk,kt->nft', st['Z'], st['T'], st['V'])


def reconstruct_covariance(st):
    """X_hat[f,t] = sum_k (sum_n H[f,n] Z[n,k]) T[f,k] V[k,t].  src/bss/mnmf.py:554-562."""
    HZ = np.einsum('fnij,nk->fkij', st['H'], st['Z'])
    return np.einsum('fkij,fk,kt->ftij', HZ, st['T'], st['V'])


def _traces(st, eps):
    """The two (F,N,T) trace tensors every multiplicative update starts from (:392-398):
    tr(X_hat^-1 X X_hat^-1 H_n) and tr(X_hat^-1 H_n), with X[f,t] = x x^H (:222-223)."""
    X, H = st['X'], st['H']
    C = X.shape[0]
    inv = np.linalg.inv(reconstruct_covariance(st) + eps * np.eye(C))        # (F,T,C,C)
    x = X.transpose(1, 2, 0)                                                 # (F,T,C)
    XX = x[..., :, None] * x[..., None, :].conj()
    XXX = inv @ XX @ inv
    num = np.einsum('ftij,fnji->fnt', XXX, H).real
    den = np.einsum('ftij,fnji->fnt', inv, H).real
    return num, den, inv, XXX


def update_basis(st, eps=EPS):
    """src/bss/mnmf.py:381-403."""
    a, b, _, _ = _traces(st, eps)
    num = np.einsum('nk,kt,fnt->fk', st['Z'], st['V'], a)
    den = np.einsum('nk,kt,fnt->fk', st['Z'], st['V'], b)
    den[den < eps] = eps
    st['T'] = st['T'] * np.sqrt(num / den)


def update_activation(st, eps=EPS):
    """src/bss/mnmf.py:405-427."""
    a, b, _, _ = _traces(st, eps)
    num = np.einsum('nk,fk,fnt->kt', st['Z'], st['T'], a)
    den = np.einsum('nk,fk,fnt->kt', st['Z'], st['T'], b)
    den[den < eps] = eps
    st['V'] = st['V'] * np.sqrt(num / den)


def update_latent(st, eps=EPS):
    """src/bss/mnmf.py:429-453 (columns of Z renormalised to sum to one)."""
    a, b, _, _ = _traces(st, eps)
    num = np.einsum('fk,kt,fnt->nk', st['T'], st['V'], a)
    den = np.einsum('fk,kt,fnt->nk', st['T'], st['V'], b)
    den[den < eps] = eps
    Z = st['Z'] * np.sqrt(num / den)
    s = Z.sum(axis=0)
    s[s < eps] = eps
    st['Z'] = Z / s


def solve_riccati(A, B):
    """H with H A H = B from the stable invariant subspace of [[0,-A],[-B,0]].  src/algorithm/linalg.py:7-30:
    eigenvectors of the M smallest real parts, H = G F^-1, Hermitian part."""
    M = A.shape[-1]
    O = np.zeros_like(A)
    L = np.concatenate([np.concatenate([O, -A], axis=-1), np.concatenate([-B, O], axis=-1)], axis=-2)
    w, v = np.linalg.eig(L)
    order = np.argsort(np.real(w), axis=-1)[..., :M]
    FG = np.take_along_axis(v, order[..., None, :], axis=-1)                # columns of the chosen eigenvectors
    Fm, Gm = FG[..., :M, :], FG[..., M:, :]
    Hs = Gm @ np.linalg.inv(Fm)
    return (Hs + Hs.swapaxes(-1, -2).conj()) / 2


def solve_riccati_hermitian(A, B):
    """The same solution in closed form, H = A^-1/2 (A^1/2 B A^1/2)^1/2 A^-1/2 (what the CUDA path computes):
    two Hermitian eigen-decompositions instead of a 2M x 2M non-symmetric one."""
    def fn(M, f):
        w, v = np.linalg.eigh((M + M.swapaxes(-1, -2).conj()) / 2)
        return (v * f(np.maximum(w, 0))[..., None, :]) @ v.swapaxes(-1, -2).conj()
    S = fn(A, np.sqrt)
    Si = fn(A, lambda w: 1 / np.sqrt(w))
    Hs = Si @ fn(S @ B @ S, np.sqrt) @ Si
    return (Hs + Hs.swapaxes(-1, -2).conj()) / 2


def update_spatial(st, normalize=True, eps=EPS, riccati=solve_riccati):
    """src/bss/mnmf.py:455-483."""
    H = st['H']
    C = H.shape[-1]
    _, _, inv, XXX = _traces(st, eps)
    lam = source_power(st)                                                   # ZT (:472) contracted with V (:470-471)
    A = np.einsum('nft,ftij->fnij', lam, inv)
    Bm = H @ np.einsum('nft,ftij->fnij', lam, XXX) @ H
    H = riccati(A, Bm) + eps * np.eye(C)
    if normalize:
        H = H / np.trace(H, axis1=2, axis2=3)[..., None, None]
    st['H'] = H


def update_once(st, normalize=True, eps=EPS, riccati=solve_riccati):
    """src/bss/mnmf.py:311-315."""
    update_basis(st, eps)
    update_activation(st, eps)
    update_latent(st, eps)
    update_spatial(st, normalize, eps, riccati)


def to_psd(X, eps=EPS):
    """src/utils/utils_linalg.py:9-31."""
    C = X.shape[-1]
    X = (X + X.swapaxes(-1, -2).conj()) / 2
    delta = np.minimum(np.linalg.eigvalsh(X).min(axis=-1), 0)
    trace = np.trace(X, axis1=-2, axis2=-1).real
    return X - delta[..., None, None] * np.eye(C) + eps * trace[..., None, None] * np.eye(C)


def logdet_divergence(inp, target, eps=EPS):
    """src/criterion/divergence.py:83-106."""
    C = inp.shape[-1]
    trace = np.trace(target @ np.linalg.inv(inp), axis1=-2, axis2=-1).real
    ex, ey = np.linalg.eigvalsh(target).real, np.linalg.eigvalsh(inp).real
    ex[ex < eps], ey[ey < eps] = eps, eps
    return trace - (np.sum(np.log(ex), axis=-1) - np.sum(np.log(ey), axis=-1)) - C


def negative_loglikelihood(st, eps=EPS):
    """src/bss/mnmf.py:575-589."""
    X = st['X']
    C = X.shape[0]
    x = X.transpose(1, 2, 0)
    XX = x[..., :, None] * x[..., None, :].conj()
    A = to_psd(XX) + eps * np.eye(C)
    Bm = to_psd(reconstruct_covariance(st)) + eps * np.eye(C)
    return logdet_divergence(Bm, A).sum()


def separate(st, reference_id=0, eps=EPS):
    """Multichannel Wiener filter, image at the reference microphone (N,F,T).  src/bss/mnmf.py:609-634."""
    X, H = st['X'], st['H']
    C = X.shape[0]
    inv = np.linalg.inv(reconstruct_covariance(st) + eps * np.eye(C))
    q = np.einsum('ftij,jft->fti', inv, X)
    Hq = np.einsum('fnj,ftj->nft', H[:, :, reference_id, :], q)
    return source_power(st) * Hq


def run(X, iteration=100, n_basis=10, n_sources=None, normalize=True, reference_id=0, eps=EPS, record_loss=True,
        riccati=solve_riccati, **presets):
    """src/bss/mnmf.py:153-184."""
    st = init_state(X, n_basis, n_sources, eps=eps, **presets)
    loss = [negative_loglikelihood(st, eps)] if record_loss else None
    for _ in range(iteration):
        update_once(st, normalize, eps, riccati)
        if record_loss:
            loss.append(negative_loglikelihood(st, eps))
    return separate(st, reference_id, eps), st, loss
