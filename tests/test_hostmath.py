"""CPU: the per-bin fp64 linear algebra the kernels run (csrc/smallmat.cuh, compiled for the host by
tests/hostmath/hostmath.cpp) against the oracle's NumPy/LAPACK formulation."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, load_golden, rel
from oracle import core

SRC = os.path.join(ROOT, 'tests', 'hostmath', 'hostmath.cpp')
OUT = os.path.join(ROOT, 'tests', 'hostmath', '_hostmath.so')


@pytest.fixture(scope='module')
def hm():
    hdr = os.path.join(ROOT, 'audio_source_separation_b200', 'csrc', 'smallmat.cuh')
    if not os.path.exists(OUT) or os.path.getmtime(OUT) < max(os.path.getmtime(SRC), os.path.getmtime(hdr)):
        subprocess.check_call(['g++', '-O2', '-std=c++17', '-shared', '-fPIC', '-x', 'c++', SRC, '-o', OUT])
    lib = ctypes.CDLL(OUT)
    vp, i, d = ctypes.c_void_p, ctypes.c_int, ctypes.c_double
    lib.hm_ip_sweep.argtypes = [i, i, vp, vp, vp, d, i, i, d]
    lib.hm_cond2.argtypes = [i, vp, vp]
    lib.hm_inverse.argtypes = [i, vp, vp]
    lib.hm_det.argtypes = [i, vp, vp]
    lib.hm_riccati.argtypes = [i, vp, vp, vp]
    lib.hm_eigh.argtypes = [i, vp, vp, vp]
    return lib


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _rand_c(rng, *shape):
    return rng.standard_normal(shape) + 1j * rng.standard_normal(shape)


@pytest.mark.parametrize('C', [2, 3, 4, 5, 8])
def test_inverse_det_cond(hm, C):
    rng = np.random.default_rng(C)
    for trial in range(20):
        A = np.ascontiguousarray(_rand_c(rng, C, C))
        inv = np.empty_like(A)
        assert hm.hm_inverse(C, _p(A), _p(inv)) == 1
        assert rel(inv, np.linalg.inv(A)) < 1e-10
        det = np.empty(2)
        hm.hm_det(C, _p(A), _p(det))
        assert abs(complex(det[0], det[1]) - np.linalg.det(A)) < 1e-10 * abs(np.linalg.det(A))
        c = np.empty(1)
        hm.hm_cond2(C, _p(A), _p(c))
        assert abs(c[0] - np.linalg.cond(A)) < 1e-8 * np.linalg.cond(A)


def test_cond_of_ill_conditioned_matrix(hm):
    """The one-sided Jacobi must stay accurate where an eigen-decomposition of A^H A would not."""
    rng = np.random.default_rng(0)
    C = 4
    for target in (1e6, 1e10, 0.7e12, 1.5e12, 1e14):
        Q1, _ = np.linalg.qr(_rand_c(rng, C, C))
        Q2, _ = np.linalg.qr(_rand_c(rng, C, C))
        s = np.array([1.0, 0.3, 0.01, 1.0 / target])
        A = np.ascontiguousarray((Q1 * s) @ Q2.conj().T)
        c = np.empty(1)
        hm.hm_cond2(C, _p(A), _p(c))
        want = np.linalg.cond(A)
        # LAPACK's own relative accuracy on sigma_min is ~eps * cond: compare within that
        assert abs(c[0] - want) < max(1e-3, 4e-16 * target) * want
        assert abs(c[0] - target) < max(1e-3, 4e-16 * target) * target


def test_singular_matrix_is_flagged(hm):
    A = np.zeros((3, 3), dtype=np.complex128)
    A[0, 0] = 1
    inv = np.empty_like(A)
    assert hm.hm_inverse(3, _p(A), _p(inv)) == 0


def _sweep(hm, W, U, thr=1e12, use_gate=True, floor_den=False, eps=1e-12):
    F, N, C = W.shape
    W = np.array(W, dtype=np.complex128, order='C', copy=True)
    U = np.ascontiguousarray(U, dtype=np.complex128)
    gate = np.empty((N, F), dtype=np.int32)
    n_sing = hm.hm_ip_sweep(C, F, _p(W), _p(U), _p(gate), thr, int(use_gate), int(floor_den), eps)
    return W, gate.astype(bool), n_sing


@pytest.mark.parametrize('C', [2, 3, 4, 6])
def test_ip_sweep_matches_oracle(hm, C):
    from oracle import synth
    F, T = 11, 40
    X = synth.mix2(C, F, T, seed=C)
    rng = np.random.default_rng(1)
    R = 10 ** rng.uniform(-3, 1, size=(C, F, T))
    U = core.weighted_covariance(X, R)
    W0 = synth.random_demix(C, F, seed=5)
    for floor in (False, True):
        Wo = W0.copy()
        gate_o = core.ip_rows(Wo, U, den_floor=1e-12 if floor else None)
        Wg, gate_g, n_sing = _sweep(hm, W0, U, floor_den=floor)
        assert n_sing == 0
        assert np.array_equal(gate_g, gate_o)
        assert rel(Wg, Wo) < 1e-10


def test_ip_gate_decisions_near_threshold(hm):
    """Purpose-built near-singular bins: the gate must take the reference's decision on both sides of 1e12."""
    rng = np.random.default_rng(3)
    C, F = 3, 24
    U = np.empty((C, F, C, C), dtype=np.complex128)
    conds = 10 ** np.linspace(9, 15, F)
    for n in range(C):
        for f in range(F):
            Q, _ = np.linalg.qr(_rand_c(rng, C, C))
            s = np.array([1.0, 0.1, 1.0 / conds[f]])
            U[n, f] = (Q * s) @ Q.conj().T
    W0 = np.tile(np.eye(C, dtype=np.complex128), (F, 1, 1))
    Wo = W0.copy()
    gate_o = core.ip_rows(Wo, U)
    Wg, gate_g, _ = _sweep(hm, W0, U)
    # only the first row sees the designed matrices unchanged (W = I): its gate follows cond(U_0)
    want = np.linalg.cond(U[0]) < 1e12
    assert np.array_equal(gate_g[0], want)
    assert np.array_equal(gate_g[0], gate_o[0])
    assert gate_g[0].any() and not gate_g[0].all()
    ok = gate_o[0]
    assert rel(Wg[ok, 0], Wo[ok, 0]) < 1e-3   # cond up to 1e12: eps * cond
    assert np.array_equal(Wg[~ok, 0], W0[~ok, 0])


@pytest.mark.parametrize('C', [2, 3, 4, 5, 8])
def test_ip_gate_sweep_of_condition_numbers(hm, C):
    """The gate takes a determinant-based shortcut when cond_2 is certainly below the threshold and the exact route
    otherwise: over twelve decades of condition numbers (dense around 1e12, several singular-value profiles) the decision
    must be numpy's `cond(W U) < 1e12`, and the updated rows must agree with the oracle wherever the gate passes."""
    rng = np.random.default_rng(40 + C)
    exps = np.concatenate((np.linspace(0, 11, 23), np.linspace(11.0, 13.0, 41), np.linspace(13, 15, 5)))
    F = len(exps)
    U = np.empty((C, F, C, C), dtype=np.complex128)
    for n in range(C):
        for f in range(F):
            Q, _ = np.linalg.qr(_rand_c(rng, C, C))
            lo = 10.0 ** -exps[f]
            if n % 3 == 0:       # one small singular value
                sv = np.concatenate((np.ones(C - 1), [lo]))
            elif n % 3 == 1:     # geometric spread
                sv = lo ** (np.arange(C) / (C - 1))
            else:                # one large singular value
                sv = np.concatenate(([1.0], np.full(C - 1, lo)))
            U[n, f] = (Q * (sv * 10.0 ** rng.uniform(-3, 3))) @ Q.conj().T
    W0 = np.tile(np.eye(C, dtype=np.complex128), (F, 1, 1))
    Wg, gate_g, n_sing = _sweep(hm, W0, U)
    assert n_sing == 0
    # row 0 sees W = I, so its matrix is U[0] itself; the later rows see whatever the earlier updates left
    cond0 = np.linalg.cond(U[0])
    clear = np.abs(np.log10(cond0) - 12.0) > 1e-3      # numpy's own SVD is only good to ~1e-4 relative at cond 1e12
    assert np.array_equal(gate_g[0][clear], (cond0 < 1e12)[clear])
    Wo = W0.copy()
    gate_o = core.ip_rows(Wo, U)
    agree = gate_g == gate_o
    assert agree[0][clear].all()
    ok = gate_o[0] & gate_g[0] & (cond0 < 1e9)
    assert rel(Wg[ok, 0], Wo[ok, 0]) < 1e-6


@pytest.mark.parametrize('C', [2, 3, 4])
def test_hermitian_eig_and_riccati(hm, C):
    """Jacobi eigen-decomposition and the closed-form Riccati solution the IS-MNMF spatial update runs, against
    LAPACK `eigh` and the oracle's restatement of solve_Riccati (src/algorithm/linalg.py:7-30)."""
    from oracle import mnmf
    rng = np.random.default_rng(10 + C)
    for trial in range(20):
        G = _rand_c(rng, C, C)
        A = np.ascontiguousarray(G @ G.conj().T + 0.1 * np.eye(C))
        G = _rand_c(rng, C, C)
        B = np.ascontiguousarray(G @ G.conj().T * 10.0 ** rng.integers(-3, 3))
        vecs, vals = np.empty_like(A), np.empty(C)
        hm.hm_eigh(C, _p(A), _p(vecs), _p(vals))
        assert rel(np.sort(vals), np.linalg.eigvalsh(A)) < 1e-12
        assert rel((vecs * vals) @ vecs.conj().T, A) < 1e-12
        H = np.empty_like(A)
        hm.hm_riccati(C, _p(A), _p(B), _p(H))
        assert rel(H @ A @ H, B) < 1e-10
        assert rel(H, mnmf.solve_riccati(A[None, None], B[None, None])[0, 0]) < 1e-9
    # identity / diagonal inputs (the very first spatial update starts from H = I)
    A = np.ascontiguousarray(np.diag(rng.random(C) + 0.5).astype(np.complex128))
    B = np.ascontiguousarray(np.diag(rng.random(C) + 0.5).astype(np.complex128))
    H = np.empty_like(A)
    hm.hm_riccati(C, _p(A), _p(B), _p(H))
    assert rel(H, np.diag(np.sqrt(np.diag(B).real / np.diag(A).real))) < 1e-13
