"""GPU parity of the single-channel NMF multiplicative updates (src/algorithm/nmf.py) against the golden fixtures
generated from the unmodified reference and against the oracle.  The kernels run in float64, so the tolerance is
1e-9 relative on the factors and on every entry of the loss history.
"""
import numpy as np
import pytest

from conftest import load_golden, rel
from oracle import nmf as o_nmf, synth

pytestmark = pytest.mark.gpu

NMF_CASES = ['nmf_euc_d2', 'nmf_euc_d1', 'nmf_kl_d2', 'nmf_kl_d1', 'nmf_is_mm_d2', 'nmf_is_mm_d1', 'nmf_is_me', 'nmf_t',
             'nmf_cauchy_naive_multipricative', 'nmf_cauchy_mm', 'nmf_cauchy_me', 'nmf_cauchy_mm_fast']
TOL = 1e-9


def _make(meta):
    from audio_source_separation_b200.algorithm.nmf import EUCNMF, KLNMF, ISNMF, tNMF, CauchyNMF
    kind = meta['model']
    if kind == 't':
        return tNMF(n_basis=meta['n_basis'], nu=meta['nu'], domain=meta['domain'], algorithm=meta['algorithm'])
    cls = {'euc': EUCNMF, 'kl': KLNMF, 'is': ISNMF, 'cauchy': CauchyNMF}[kind]
    return cls(n_basis=meta['n_basis'], domain=meta['domain'], algorithm=meta['algorithm'])


@pytest.mark.parametrize('name', NMF_CASES)
def test_nmf_golden_seeded_dropin(cuda_device, name):
    """Seeded exactly like the reference run that produced the fixture: `_reset` must consume the global RNG in
    the reference's order (basis, then activation)."""
    meta, i, o = load_golden(name)
    model = _make(meta)
    np.random.seed(meta['seed'])
    T, V = model(i['Z'], iteration=meta['iteration'])
    assert T.shape == o['basis'].shape and V.shape == o['activation'].shape
    assert rel(T, o['basis']) < TOL
    assert rel(V, o['activation']) < TOL
    assert len(model.loss) == meta['iteration']
    assert np.max(np.abs(np.array(model.loss) - o['loss']) / np.abs(o['loss'])) < TOL


def test_nmf_update_once_by_hand(cuda_device):
    """update_once driven by the user with an injected state; host assignments between iterations are honoured."""
    from audio_source_separation_b200.algorithm.nmf import ISNMF
    meta, i, o = load_golden('nmf_is_mm_d2')
    model = ISNMF(n_basis=meta['n_basis'])
    model.target = i['Z']
    model._reset()
    model.basis, model.activation = i['T0'].copy(), i['V0'].copy()
    T, V = i['T0'].copy(), i['V0'].copy()
    for _ in range(3):
        model.update_once()
        T, V = o_nmf.is_step(i['Z'], T, V)
        assert rel(model.basis, T) < TOL and rel(model.activation, V) < TOL
    model.activation = np.asarray(model.activation) * 2.0
    V = V * 2.0
    model.update_once()
    T, V = o_nmf.is_step(i['Z'], T, V)
    assert rel(model.basis, T) < TOL and rel(model.activation, V) < TOL


@pytest.mark.parametrize('F,T,K', [(257, 128, 4), (33, 1000, 11), (600, 37, 1), (5, 3, 20)])
def test_nmf_shapes_vs_oracle(cuda_device, F, T, K):
    """cfg1 shape (257 x 128, K = 4) and ragged shapes: K above/below the accumulation chunk, T not a multiple of the
    block, more basis vectors than bins."""
    from audio_source_separation_b200.algorithm.nmf import EUCNMF, KLNMF
    Z = synth.spectrogram(F, T, seed=F + T)
    for cls, kind in ((EUCNMF, 'euc'), (KLNMF, 'kl')):
        np.random.seed(3)
        model = cls(n_basis=K, domain=1.5)
        Tg, Vg = model(Z, iteration=5)
        np.random.seed(3)
        To, Vo, lo = o_nmf.run(kind, Z, n_basis=K, iteration=5, domain=1.5)
        assert rel(Tg, To) < TOL and rel(Vg, Vo) < TOL
        assert np.allclose(model.loss, lo, rtol=1e-9, atol=0)


def test_nmf_loss_decreases_and_floor(cuda_device):
    """EUC-NMF loss is non-increasing (MM algorithm); an all-zero target exercises the eps floors."""
    from audio_source_separation_b200.algorithm.nmf import EUCNMF
    Z = synth.spectrogram(257, 128, seed=0)
    np.random.seed(111)
    model = EUCNMF(n_basis=4)
    model(Z, iteration=50)
    assert all(b <= a * (1 + 1e-12) for a, b in zip(model.loss, model.loss[1:]))
    Z0 = np.zeros((9, 7))
    np.random.seed(1)
    T, V = EUCNMF(n_basis=2)(Z0, iteration=3)
    np.random.seed(1)
    To, Vo, _ = o_nmf.run('euc', Z0, n_basis=2, iteration=3)
    assert np.allclose(T, To, rtol=1e-9, atol=1e-300) and np.allclose(V, Vo, rtol=1e-9, atol=1e-300)


def test_nmf_argument_errors(cuda_device):
    from audio_source_separation_b200.algorithm.nmf import EUCNMF, ISNMF, CauchyNMF
    with pytest.raises(AssertionError):
        EUCNMF(domain=3)
    with pytest.raises(AssertionError):
        EUCNMF(algorithm='me')
    with pytest.raises(AssertionError):
        CauchyNMF(n_basis=2, domain=1)
    with pytest.raises(AssertionError):
        EUCNMF().update_once()            # "Specify data!"
    Z = synth.spectrogram(9, 7, seed=0)
    with pytest.raises(ValueError):
        ISNMF(algorithm='nope')(Z, iteration=1)
    with pytest.raises(AssertionError):
        ISNMF(algorithm='me', domain=1)(Z, iteration=1)


@pytest.mark.parametrize('cls_name,kw', [('EUCNMF', {}), ('KLNMF', dict(domain=1.5)), ('ISNMF', dict(algorithm='me')), ('tNMF', dict(nu=50.0)),
                                         ('CauchyNMF', dict(algorithm='mm_fast'))])
def test_nmf_cluster_kernel_equals_phase_kernels(cuda_device, monkeypatch, cls_name, kw):
    """Small problems run as one thread-block-cluster launch for the whole loop (target rows split over 8 CTAs, partial
    sums exchanged through distributed shared memory); it must agree with the per-phase kernels it replaces, loss history
    included, for a row count that is not a multiple of the cluster size."""
    import audio_source_separation_b200.algorithm.nmf as nmf_mod
    Z = synth.spectrogram(61, 97, seed=11)
    results = []
    for no_cluster in (False, True):
        if no_cluster:
            monkeypatch.setenv('BSSGPU_NO_CLUSTER', '1')
        else:
            monkeypatch.delenv('BSSGPU_NO_CLUSTER', raising=False)
        np.random.seed(5)
        model = getattr(nmf_mod, cls_name)(n_basis=3, **kw)
        T, V = model(Z, iteration=7)
        model.update_once()                       # single update through the same path
        results.append((T, V, np.array(model.loss), np.asarray(model.basis)))
    for a, b in zip(results[0], results[1]):
        assert np.allclose(a, b, rtol=1e-11, atol=0)
