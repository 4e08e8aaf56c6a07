"""Edge cases of the update loop against the oracle: ragged shapes (odd / tiny frame counts, frame counts around the
128-frame block and the 64-frame lane stride, a single bin), every channel count, n_basis above and below the kernels'
compile-time fast path, non-integer domain, re-use of a model on new inputs."""
import warnings

import numpy as np
import pytest

from conftest import rel
from oracle import ilrma as o_ilrma, auxiva as o_auxiva, fastmnmf as o_mnmf, synth

pytestmark = pytest.mark.gpu

TOL = 3e-4


def _ilrma_pair(X, K, n_iter, seed=7, **kw):
    from audio_source_separation_b200.bss.ilrma import GaussILRMA
    C, F, T = X.shape
    W0, T0, V0 = synth.initial_state(C, F, T, K, seed=seed)
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        model = GaussILRMA(n_basis=K, **kw)
        out = model(X, iteration=n_iter, demix_filter=W0, basis=T0, activation=V0)
    spatial = kw.get('algorithm_spatial', 'IP')
    want, st, loss = o_ilrma.run(X, iteration=n_iter, n_basis=K, spatial=spatial, domain=kw.get('domain', 2),
                                 normalize_mode=kw.get('normalize', 'power'), W=W0, T=T0, V=V0)
    return model, out, want, st, loss


@pytest.mark.parametrize('C,F,T,K', [(2, 1, 40, 2), (2, 5, 3, 2), (3, 7, 63, 2), (3, 7, 65, 3), (4, 9, 127, 2), (4, 9, 129, 2),
                                     (4, 6, 257, 10), (5, 4, 130, 2), (6, 3, 70, 4), (7, 3, 50, 2), (8, 3, 140, 2), (4, 300, 16, 2)])
def test_ilrma_ip_ragged_shapes(cuda_device, C, F, T, K):
    X = synth.mix2(C, F, T, seed=C * 100 + T)
    model, out, want, st, loss = _ilrma_pair(X, K, 2)
    assert rel(out, want) < TOL
    assert rel(model.basis, st['T']) < TOL and rel(model.activation, st['V']) < TOL and rel(model.demix_filter, st['W']) < TOL
    assert np.max(np.abs(np.array(model.loss) - np.array(loss)) / np.maximum(np.abs(loss), 1e-3 * np.max(np.abs(loss)))) < 2e-4


@pytest.mark.parametrize('kw', [dict(algorithm_spatial='ISS'), dict(algorithm_spatial='IP2'), dict(domain=1.5),
                                dict(normalize='projection-back', domain=1.2), dict(algorithm_spatial='ISS', normalize='projection-back')])
def test_ilrma_variants_odd_frames(cuda_device, kw):
    X = synth.mix2(3, 11, 131, seed=5)
    model, out, want, st, loss = _ilrma_pair(X, 3, 2, **kw)
    assert rel(out, want) < TOL
    assert rel(model.basis, st['T']) < TOL and rel(model.activation, st['V']) < TOL


@pytest.mark.parametrize('C,F,T,spatial,kind', [(2, 9, 3, 'IP', 'laplace'), (5, 6, 129, 'IP', 'gauss'), (8, 4, 65, 'IP', 'laplace'),
                                                (4, 7, 131, 'ISS', 'laplace'), (6, 5, 77, 'ISS', 'gauss'), (4, 9, 130, 'IP2', 'laplace')])
def test_auxiva_ragged_shapes(cuda_device, C, F, T, spatial, kind):
    from audio_source_separation_b200.bss.iva import AuxLaplaceIVA, AuxGaussIVA
    X = synth.mix2(C, F, T, seed=C + T)
    cls = AuxLaplaceIVA if kind == 'laplace' else AuxGaussIVA
    model = cls(algorithm_spatial=spatial)
    out = model(X, iteration=2)
    want, st, loss = o_auxiva.run(X, iteration=2, kind=kind, spatial=spatial)
    assert rel(out, want) < TOL
    assert np.max(np.abs(np.array(model.loss) - np.array(loss)) / np.maximum(np.abs(loss), 1e-3 * np.max(np.abs(loss)))) < 2e-4


@pytest.mark.parametrize('M,N,F,T,K', [(2, 2, 5, 3, 2), (4, 4, 6, 129, 2), (3, 3, 4, 131, 1), (6, 2, 3, 66, 3)])
def test_fastmnmf_ragged_shapes(cuda_device, M, N, F, T, K):
    from audio_source_separation_b200.bss.mnmf import FastMultichannelISNMF
    X = synth.mix2(M, F, T, seed=M + N + T)
    rng = np.random.default_rng(1)
    W0 = rng.random((N, F, K)).astype(np.float32).astype(np.float64)
    H0 = rng.random((N, K, T)).astype(np.float32).astype(np.float64)
    model = FastMultichannelISNMF(n_basis=K, n_sources=N)
    out = model(X, iteration=2, basis=W0, activation=H0)
    want, st, loss = o_mnmf.run(X, iteration=2, n_basis=K, n_sources=N, W=W0, H=H0)
    assert rel(out, want) < TOL
    assert rel(model.diagonalizer, st['Q']) < TOL and rel(model.spatial_covariance, st['G']) < TOL
    assert np.max(np.abs(np.array(model.loss) - np.array(loss)) / np.abs(loss)) < 2e-4


def test_model_reuse_on_new_inputs(cuda_device):
    """One model object called on a second mixture of the same shape and then of another shape (handle re-creation):
    no state may leak from the previous problem."""
    from audio_source_separation_b200.bss.ilrma import GaussILRMA
    model = GaussILRMA(n_basis=2, recordable_loss=False)
    for C, F, T, seed in ((3, 17, 40, 1), (3, 17, 40, 2), (2, 9, 70, 3)):
        X = synth.mix2(C, F, T, seed=seed)
        W0, T0, V0 = synth.initial_state(C, F, T, 2, seed=seed)
        out = model(X, iteration=3, demix_filter=W0, basis=T0, activation=V0)
        want, _, _ = o_ilrma.run(X, iteration=3, n_basis=2, W=W0, T=T0, V=V0, record_loss=False)
        assert rel(out, want) < TOL


def test_tilrma_shapes(cuda_device):
    from audio_source_separation_b200.bss.ilrma import tILRMA
    for C, F, T, K, nu in ((2, 9, 65, 2, 1.0), (4, 5, 130, 3, 100.0)):
        X = synth.mix2(C, F, T, seed=C + F)
        W0, T0, V0 = synth.initial_state(C, F, T, K, seed=3)
        model = tILRMA(n_basis=K, nu=nu)
        out = model(X, iteration=2, demix_filter=W0, basis=T0, activation=V0)
        want, st, loss = o_ilrma.t_run(X, iteration=2, n_basis=K, nu=nu, W=W0, T=T0, V=V0)
        assert rel(out, want) < TOL and rel(model.basis, st['T']) < TOL
