"""Shared building blocks of the oracle (test infrastructure, see oracle/__init__.py).

Array conventions follow the reference: X (C,F,T) mixture, W (F,N,C) demixing
filters, Y (N,F,T) estimates, U (N,F,C,C) weighted covariances.  Everything is
float64 / complex128.
"""
import numpy as np

EPS = 1e-12        # src/bss/ilrma.py:8, src/bss/iva.py:8, src/bss/mnmf.py:9
THRESHOLD = 1e12   # src/bss/ilrma.py:9, src/bss/iva.py:9, src/bss/mnmf.py:10


def solve_vec(A, b):
    """Stacked `A x = b` with a stack of right-hand-side *vectors* b (...,C).

    The reference calls np.linalg.solve(A (F,C,C), b (F,C)) (src/bss/ilrma.py:523)
    which under NumPy 1.x (the reference's era) means exactly this; NumPy >= 2
    changed the broadcasting rule, so the vector axis is made explicit here.
    """
    return np.linalg.solve(A, b[..., np.newaxis])[..., 0]


def demix(X, W):
    """Y[n,f,t] = sum_c W[f,n,c] X[c,f,t].  src/bss/ilrma.py:153-165, src/bss/iva.py:105-117."""
    return (W @ X.transpose(1, 0, 2)).transpose(1, 0, 2)


def estimate_demix_filter(Y, X):
    """Least-squares W with Y = W X per bin.  src/bss/ilrma.py:167-173, src/bss/iva.py:119-125."""
    Xh = X.transpose(1, 2, 0).conj()                  # (F,T,C)
    G = X.transpose(1, 0, 2) @ Xh                     # (F,C,C)
    return Y.transpose(1, 0, 2) @ Xh @ np.linalg.inv(G)


def projection_back_scale(Y, reference):
    """Per-bin least-squares scale.  src/algorithm/projection_back.py:3-34.

    reference (F,T)  -> scale (N,F)
    reference (C,F,T) -> scale (C,N,F)
    """
    if reference.ndim == 2:
        ref = reference[np.newaxis]
    elif reference.ndim == 3:
        ref = reference
    else:
        raise ValueError("reference.ndim is expected 2 or 3, but given {}.".format(reference.ndim))
    Xb = ref.transpose(1, 0, 2)                       # (F,Cr,T)
    Yb = Y.transpose(1, 0, 2)                         # (F,N,T)
    Yh = Yb.transpose(0, 2, 1).conj()                 # (F,T,N)
    A = Xb @ Yh @ np.linalg.inv(Yb @ Yh)              # (F,Cr,N)
    if reference.ndim == 2:
        return A[:, 0, :].transpose(1, 0)
    return A.transpose(1, 2, 0)


def gather_by_order(x, order, axis=-2):
    """Batched gather along `axis` by integer `order`.  src/utils/utils_linalg.py:33-52.

    Pure index arithmetic: results must be bit-identical to the reference.
    """
    lead = x.shape[:axis]
    n_elem = x.shape[axis]
    tail = x.shape[axis + 1:]
    n_pick = order.shape[-1]
    flat = x.reshape(-1, *tail)
    base = np.repeat(n_elem * np.arange(int(np.prod(lead))), n_pick)
    picked = flat[order.reshape(-1) + base]
    return picked.reshape(*lead, n_pick, *tail)


def next_update_pair(pair, n_sources):
    """IP2 pair schedule.  src/bss/ilrma.py:635-646, src/bss/iva.py:372-383."""
    if pair is None:
        return (0, 1)
    m, n = pair
    return ((m + 1) % n_sources, (n + 1) % n_sources)


def outer_products(X):
    """XX[f,t] = x_ft x_ft^H, (F,T,C,C).  src/bss/ilrma.py:505-508."""
    Xc = X.transpose(1, 2, 0)[..., np.newaxis]        # (F,T,C,1)
    return Xc @ Xc.transpose(0, 1, 3, 2).conj()


def weighted_covariance(X, R):
    """U[n,f] = mean_t x x^H / R[n,f,t]  with R broadcastable to (N,F,T).

    src/bss/ilrma.py:503-511 (R = (TV)^(2/d), already floored by the caller),
    src/bss/iva.py:491-499 and :724-732 (R = r[n,1,t]),
    src/bss/mnmf.py:875 (one channel of R[f,t,m] at a time).
    Materialises XX/R exactly like the reference (this is what the CPU baseline times).
    """
    XX = outer_products(X)                            # (F,T,C,C)
    return (XX / R[..., np.newaxis, np.newaxis]).mean(axis=-3)


def ip_rows(W, U, threshold=THRESHOLD, den_floor=None):
    """Gauss-Seidel iterative-projection row updates, W modified in place.

    src/bss/ilrma.py:512-530, src/bss/iva.py:500-518 and :733-751 (no floor on
    the denominator); src/bss/mnmf.py:872-886 passes den_floor=eps (:883).
    Returns the (N,F) boolean gate mask `cond(WU) < threshold`.
    """
    n_rows = W.shape[1]
    n_bins, n_ch = W.shape[0], W.shape[2]
    E = np.tile(np.eye(n_rows, n_ch), (n_bins, 1, 1))
    gate = np.zeros((n_rows, n_bins), dtype=bool)
    for n in range(n_rows):
        U_n = U[n]
        WU = W @ U_n
        ok = np.linalg.cond(WU) < threshold
        w = solve_vec(WU, E[:, n, :])
        q = w[:, np.newaxis, :].conj() @ U_n @ w[:, :, np.newaxis]
        den = np.sqrt(q[..., 0])                      # (F,1) complex
        if den_floor is not None:
            den[den < den_floor] = den_floor
        W[:, n, :] = np.where(ok[:, np.newaxis], w.conj() / den, W[:, n, :])
        gate[n] = ok
    return gate


def ip2_pair(W, U_m, U_n, m, n, threshold=THRESHOLD):
    """Pairwise (IP2) update of rows m and n, W modified in place.

    src/bss/ilrma.py:599-626, src/bss/iva.py:566-592.
    Returns (order (F,2) int, gate_m (F,), gate_n (F,), lam (F,2) complex) for index-parity tests: `order` indexes
    `lam`, the eigenvalues in the order LAPACK returned them.
    """
    n_bins, n_ch = W.shape[0], W.shape[2]
    E = np.zeros((n_bins, n_ch, 2))
    E[:, m, 0] = 1
    E[:, n, 1] = 1
    WU_m, WU_n = W @ U_m, W @ U_n
    ok_m = np.linalg.cond(WU_m) < threshold
    ok_n = np.linalg.cond(WU_n) < threshold
    G_m = np.linalg.inv(WU_m) @ E                     # (F,C,2)
    G_n = np.linalg.inv(WU_n) @ E
    V_m = G_m.transpose(0, 2, 1).conj() @ U_m @ G_m   # (F,2,2)
    V_n = G_n.transpose(0, 2, 1).conj() @ U_n @ G_n
    lam, vec = np.linalg.eig(np.linalg.inv(V_n) @ V_m)
    order = np.argsort(lam, axis=-1)[:, ::-1]
    picked = gather_by_order(vec.swapaxes(-2, -1), order=order, axis=-2)
    v_m, v_n = picked[:, 0, :], picked[:, 1, :]
    q_m = v_m[:, np.newaxis, :].conj() @ V_m @ v_m[:, :, np.newaxis]
    q_n = v_n[:, np.newaxis, :].conj() @ V_n @ v_n[:, :, np.newaxis]
    v_m = v_m / np.sqrt(q_m.squeeze(axis=-1))
    v_n = v_n / np.sqrt(q_n.squeeze(axis=-1))
    w_m = (G_m @ v_m[..., np.newaxis]).squeeze(axis=-1).conj()
    w_n = (G_n @ v_n[..., np.newaxis]).squeeze(axis=-1).conj()
    W[:, m, :] = np.where(ok_m[:, np.newaxis], w_m, W[:, m, :])
    W[:, n, :] = np.where(ok_n[:, np.newaxis], w_n, W[:, n, :])
    return order, ok_m, ok_n, lam


def iss_sweep(Y, R):
    """One ISS sweep over all sources; R broadcastable to (N,F,T), already floored.

    src/bss/ilrma.py:557-562, src/bss/iva.py:535-540 and :768-773.
    """
    for n in range(Y.shape[0]):
        u = np.sum(Y * Y[n].conj() / R, axis=2)       # (N,F)
        d = np.sum(np.abs(Y[n]) ** 2 / R, axis=2)     # (N,F)
        v = u / d
        v[n] = 1 - 1 / np.sqrt(d[n])
        Y = Y - v[:, :, np.newaxis] * Y[n]
    return Y


def logabsdet_sum(W):
    """sum_f log|det W_f|  (src/bss/ilrma.py:675, src/bss/iva.py:617)."""
    return np.sum(np.log(np.abs(np.linalg.det(W))))
