"""GPU parity of Sawada's multichannel IS-NMF (src/bss/mnmf.py:116-635, author='Sawada') against the golden fixtures
generated from the unmodified reference and against the oracle.  The device path is fp64 end to end (the mixture is
stored as complex64, which the synthetic inputs are exactly representable in), so the tolerances are tight: 1e-8 on the
state after a few updates; the separated estimate is returned through a complex64 buffer (1e-6).  The loss is compared
at 5e-6: the reference regularises the rank-one x x^H by `to_PSD` (src/utils/utils_linalg.py:9-31), whose shift `delta`
is the most negative eigenvalue LAPACK returns for a matrix whose exact smallest eigenvalue is zero, i.e. rounding noise
of order 1e-17 |x|^2 next to the eps |x|^2 = 1e-12 |x|^2 it is added to.  The device evaluates the exact-arithmetic
value (delta = 0); the difference is a constant offset of about 1e-5 per frame on a loss of about 30 per frame.
"""
import numpy as np
import pytest

from conftest import load_golden, rel
from oracle import mnmf as o_mnmf, synth

pytestmark = pytest.mark.gpu

TOL_STATE = 1e-8
TOL_OUT = 1e-6
TOL_LOSS = 5e-6

CASES = ['mnmf_sawada_c2n2', 'mnmf_sawada_c3n2', 'mnmf_sawada_c4n3_eye']


def _initial(C, N, F, T, K, seed=5):
    rng = np.random.default_rng(seed)
    A = rng.standard_normal((F, N, C, C)) + 1j * rng.standard_normal((F, N, C, C))
    H0 = A @ A.swapaxes(-1, -2).conj() + 0.5 * np.eye(C)
    H0 = H0 / np.trace(H0, axis1=-2, axis2=-1).real[..., None, None]
    H0 = (H0 + H0.swapaxes(-1, -2).conj()) / 2
    Z0 = rng.random((N, K)) + 0.5
    Z0 = Z0 / Z0.sum(axis=0)
    return H0, Z0, rng.random((F, K)) + 0.1, rng.random((K, T)) + 0.1


def _check_state(model, st, tol=TOL_STATE):
    assert rel(model.basis, st['T']) < tol
    assert rel(model.activation, st['V']) < tol
    assert rel(model.latent, st['Z']) < tol
    assert rel(model.spatial, st['H']) < tol


@pytest.mark.parametrize('name', CASES)
def test_mnmf_sawada_golden(cuda_device, name):
    from audio_source_separation_b200.bss.mnmf import MultichannelISNMF
    meta, i, o = load_golden(name)
    model = MultichannelISNMF(n_basis=meta['n_basis'], n_sources=meta['n_sources'], normalize=meta['normalize'])
    out = model(i['X'], iteration=meta['iteration'], spatial=i['H0'], latent=i['Z0'], basis=i['T0'], activation=i['V0'])
    assert out.shape == o['output'].shape and out.dtype == np.complex128
    assert rel(model.basis, o['basis']) < TOL_STATE
    assert rel(model.activation, o['activation']) < TOL_STATE
    assert rel(model.latent, o['latent']) < TOL_STATE
    assert rel(model.spatial, o['spatial']) < TOL_STATE
    assert rel(model.loss, o['loss']) < TOL_LOSS
    assert rel(out, o['output']) < TOL_OUT
    assert model.estimation is out


@pytest.mark.parametrize('C,N,F,T,K', [(2, 2, 33, 65, 2), (2, 3, 17, 130, 3), (3, 3, 9, 257, 10), (4, 4, 17, 300, 2), (4, 8, 5, 64, 12),
                                       (3, 1, 9, 40, 2)])
def test_mnmf_sawada_update_once_vs_oracle(cuda_device, C, N, F, T, K):
    """update_once by hand from an injected state: odd frame counts, several frame blocks, K above one activation pass
    (8), more / fewer sources than channels, and each of the four sub-updates through the state they leave behind."""
    from audio_source_separation_b200.bss.mnmf import MultichannelISNMF
    X = synth.mix2(C, F, T, seed=C * 10 + N)
    H0, Z0, T0, V0 = _initial(C, N, F, T, K)
    model = MultichannelISNMF(n_basis=K, n_sources=N, recordable_loss=False)
    model.input = X
    model._reset(spatial=H0, latent=Z0, basis=T0, activation=V0)
    st = o_mnmf.init_state(X, K, N, H=H0, Z=Z0, T=T0, V=V0)
    assert abs(model.compute_negative_loglikelihood() / o_mnmf.negative_loglikelihood(st) - 1) < TOL_LOSS
    assert rel(model.estimation, o_mnmf.separate(st)) < TOL_OUT
    for it in range(2):
        model.update_once()
        o_mnmf.update_once(st)
        _check_state(model, st)
        assert abs(model.compute_negative_loglikelihood() / o_mnmf.negative_loglikelihood(st) - 1) < TOL_LOSS
    assert rel(model.separate(X), o_mnmf.separate(st)) < TOL_OUT
    # the spatial covariances stay Hermitian with unit trace (normalize=True)
    H = model.spatial
    assert np.allclose(H, H.swapaxes(-1, -2).conj(), atol=0, rtol=0)
    assert np.allclose(np.trace(H, axis1=-2, axis2=-1), 1.0, atol=1e-12)


def test_mnmf_sawada_seeded_dropin(cuda_device):
    """Default initialisation from the global NumPy state, as the reference draws it (latent, basis, activation)."""
    from audio_source_separation_b200.bss.mnmf import MultichannelISNMF
    C, F, T, K = 2, 17, 48, 4
    X = synth.mix2(C, F, T, seed=9)
    np.random.seed(111)
    model = MultichannelISNMF(n_basis=K)
    out = model(X, iteration=3)
    np.random.seed(111)
    want, st, loss = o_mnmf.run(X, iteration=3, n_basis=K)
    _check_state(model, st)
    assert rel(model.loss, loss) < TOL_LOSS and len(model.loss) == 4
    assert rel(out, want) < TOL_OUT
    assert np.all(np.diff(model.loss) < 0)


def test_mnmf_sawada_callbacks_and_assignment(cuda_device):
    """Callbacks see NumPy state after _reset and after every iteration; assigning a state attribute between updates is
    honoured (the attribute protocol of the reference classes)."""
    from audio_source_separation_b200.bss.mnmf import MultichannelISNMF
    C, N, F, T, K = 3, 2, 9, 32, 3
    X = synth.mix2(C, F, T, seed=2)
    H0, Z0, T0, V0 = _initial(C, N, F, T, K)
    seen = []
    model = MultichannelISNMF(n_basis=K, n_sources=N, callbacks=lambda m: seen.append((m.basis.copy(), m.estimation.shape)))
    model(X, iteration=2, spatial=H0, latent=Z0, basis=T0, activation=V0)
    assert len(seen) == 3 and seen[0][1] == (N, F, T)
    assert rel(seen[0][0], T0) == 0
    st = o_mnmf.init_state(X, K, N, H=H0, Z=Z0, T=T0, V=V0)
    o_mnmf.update_once(st)
    assert rel(seen[1][0], st['T']) < TOL_STATE
    o_mnmf.update_once(st)
    # overwrite the activation on both sides and continue
    V1 = np.random.default_rng(1).random((K, T)) + 0.2
    model.activation = V1
    st['V'] = V1.copy()
    model.update_once()
    o_mnmf.update_once(st)
    _check_state(model, st)


def test_mnmf_sawada_errors(cuda_device):
    from audio_source_separation_b200.bss.mnmf import MultichannelISNMF
    with pytest.raises(ValueError):
        MultichannelISNMF(hoge=1)
    with pytest.raises(AssertionError):
        MultichannelISNMF(author='nobody')
    with pytest.warns(UserWarning):
        m = MultichannelISNMF(author='Ozerov')
    with pytest.raises(NotImplementedError):
        m(synth.mix2(2, 5, 8), iteration=1)
    with pytest.raises(ValueError):
        MultichannelISNMF(n_basis=2)(synth.mix2(5, 5, 8), iteration=1)   # more than 4 channels
