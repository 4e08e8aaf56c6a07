"""Oracle for the single-channel NMF multiplicative updates (test infrastructure).

Restates src/algorithm/nmf.py: EUCNMF (:150-207), KLNMF (:209-266), ISNMF
(:268-356), tNMF (:358-428), CauchyNMF (:430-595) and the per-iteration loss
(src/criterion/divergence.py).  All updates are "T first, then V with the new T".
"""
import numpy as np

from .core import EPS


def _floor(a, eps):
    a[a < eps] = eps
    return a


def euc_step(Z, T, V, domain=2, eps=EPS):
    """src/algorithm/nmf.py:182-207."""
    a, b, q = (4 - domain) / domain, (2 - domain) / domain, domain / (4 - domain)
    TV = _floor(T @ V, eps)
    den = _floor((TV ** a) @ V.T, eps)
    T = T * (((Z * TV ** b) @ V.T) / den) ** q
    TV = _floor(T @ V, eps)
    den = _floor(T.T @ (TV ** a), eps)
    V = V * ((T.T @ (Z * TV ** b)) / den) ** q
    return T, V


def kl_step(Z, T, V, domain=2, eps=EPS):
    """src/algorithm/nmf.py:241-266."""
    b, q = (2 - domain) / domain, domain / 2
    TV = _floor(T @ V, eps)
    den = _floor((TV ** b) @ V.T, eps)
    T = T * (((Z / TV) @ V.T) / den) ** q
    TV = _floor(T @ V, eps)
    den = _floor(T.T @ (TV ** b), eps)
    V = V * ((T.T @ (Z / TV)) / den) ** q
    return T, V


def is_step(Z, T, V, domain=2, algorithm='mm', eps=EPS):
    """'mm': src/algorithm/nmf.py:302-327; 'me': :329-356 (exponent 1, domain 2 only)."""
    p = (domain + 2) / domain
    if algorithm == 'mm':
        q = domain / (domain + 2)
    elif algorithm == 'me':
        assert domain == 2, "Only domain = 2 is supported."
        q = 1
    else:
        raise ValueError("Not support {} based update.".format(algorithm))
    TV = _floor(T @ V, eps)
    den = _floor((1 / TV) @ V.T, eps)
    T = T * (((Z / TV ** p) @ V.T) / den) ** q
    TV = _floor(T @ V, eps)
    den = _floor(T.T @ (1 / TV), eps)
    V = V * ((T.T @ (Z / TV ** p)) / den) ** q
    return T, V


def t_step(Z, T, V, nu=1e3, domain=2, eps=EPS):
    """src/algorithm/nmf.py:397-428."""
    assert domain == 2, "`domain` is expected 2."
    Zf = np.maximum(Z, eps)

    def stats(T, V):
        TV = _floor(T @ V, eps)
        h = 1 / (2 / ((2 + nu) * TV) + nu / ((2 + nu) * Zf))
        return h / TV ** 2, 1 / TV

    d, inv = stats(T, V)
    T = T * np.sqrt((d @ V.T) / _floor(inv @ V.T, eps))
    d, inv = stats(T, V)
    V = V * np.sqrt((T.T @ d) / _floor(T.T @ inv, eps))
    return T, V


def cauchy_step(Z, T, V, algorithm='naive-multipricative', eps=EPS):
    """src/algorithm/nmf.py:461-595 (four algorithms, domain 2)."""
    if algorithm in ('naive-multipricative', 'mm'):
        root = algorithm == 'mm'                       # :482 vs :513
        TV = _floor(T @ V, eps)
        num = np.sum(V[np.newaxis] / TV[:, np.newaxis, :], axis=2)
        Cc = _floor(2 * Z + TV ** 2, eps)
        den = _floor(3 * (TV / Cc) @ V.T, eps)
        T = T * (np.sqrt(num / den) if root else num / den)
        TV = _floor(T @ V, eps)
        num = np.sum(T[:, :, np.newaxis] / TV[:, np.newaxis, :], axis=0)
        Cc = _floor(2 * Z + TV ** 2, eps)
        den = _floor(3 * T.T @ (TV / Cc), eps)
        V = V * (np.sqrt(num / den) if root else num / den)
        return T, V
    if algorithm == 'me':                              # :527-558 (TV itself is not floored)
        TV = T @ V
        S = _floor(TV ** 2 + Z, eps)
        A = (3 / 4) * (TV / S) @ V.T
        B = np.sum(V[np.newaxis] / TV[:, np.newaxis, :], axis=2)
        T = T * (B / _floor(A + np.sqrt(A ** 2 + 2 * B * A), eps))
        TV = T @ V
        S = _floor(TV ** 2 + Z, eps)
        A = (3 / 4) * T.T @ (TV / S)
        B = np.sum(T[:, :, np.newaxis] / TV[:, np.newaxis, :], axis=0)
        V = V * (B / _floor(A + np.sqrt(A ** 2 + 2 * B * A), eps))
        return T, V
    if algorithm == 'mm_fast':                         # :560-595
        def parts(T, V):
            TV = T @ V
            Cc = 2 * Z + TV ** 2
            CTV = _floor(Cc * TV, eps)
            ZC = Z / CTV
            Cc = _floor(Cc, eps)
            return ZC, TV / Cc
        ZC, TVC = parts(T, V)
        T = T * np.sqrt((ZC @ V.T) / _floor(TVC @ V.T, eps))
        ZC, TVC = parts(T, V)
        V = V * np.sqrt((T.T @ ZC) / _floor(T.T @ TVC, eps))
        return T, V
    raise ValueError("Not support {} based update.".format(algorithm))


def loss_value(kind, Z, T, V, domain=2, nu=1e3, eps=EPS):
    """Per-iteration loss appended by `update`.  EUC: src/algorithm/nmf.py:163,172-174;
    KL: :231-233 with src/criterion/divergence.py:34-45; IS: :290-292 with
    divergence.py:21-32; t: :367-371,387-389; Cauchy: :434-441 via NMFbase.update :51-53
    (no 2/domain power)."""
    if kind == 'cauchy':
        A = T @ V
    else:
        A = (T @ V) ** (2 / domain)
    if kind == 'euc':
        return ((Z - A) ** 2).sum()
    a, z = A + eps, Z + eps
    if kind == 'kl':
        return (z * np.log(z / a) + a - z).sum()
    if kind == 'is':
        r = z / a
        return (r - np.log(r) - 1).sum()
    if kind == 't':
        return (np.log(a) + (2 + nu) / 2 * np.log(1 + (2 / nu) * (z / a))).sum()
    if kind == 'cauchy':
        return (np.log(z / a) + 1.5 * np.log((2 * z ** 2 + a ** 2) / (3 * z ** 2))).sum()
    raise ValueError(kind)


def step(kind, Z, T, V, domain=2, algorithm='mm', nu=1e3, eps=EPS):
    if kind == 'euc':
        return euc_step(Z, T, V, domain, eps)
    if kind == 'kl':
        return kl_step(Z, T, V, domain, eps)
    if kind == 'is':
        return is_step(Z, T, V, domain, algorithm, eps)
    if kind == 't':
        return t_step(Z, T, V, nu, domain, eps)
    if kind == 'cauchy':
        return cauchy_step(Z, T, V, algorithm, eps)
    raise ValueError(kind)


def run(kind, Z, n_basis=2, iteration=100, domain=2, algorithm='mm', nu=1e3, eps=EPS, T=None, V=None):
    """NMFbase.__call__/_reset/update: src/algorithm/nmf.py:22-53.  When T/V are not given the
    initial factors are drawn from the global legacy RNG (basis first, :42-43)."""
    F, Tn = Z.shape
    T = np.random.rand(F, n_basis) if T is None else T.copy()
    V = np.random.rand(n_basis, Tn) if V is None else V.copy()
    loss = []
    for _ in range(iteration):
        T, V = step(kind, Z, T, V, domain, algorithm, nu, eps)
        loss.append(loss_value(kind, Z, T, V, domain, nu, eps))
    return T, V, loss
