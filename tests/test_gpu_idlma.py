"""GPU: GaussIDLMA drop-in (src/sss/idlma.py:88-258) against the oracle and the fixtures generated from the reference.

The DNN is the deterministic torch module of oracle/synth.py; everything else (weighted covariances from the DNN's
variances, gated IP sweep, projection-back normalisation, loss, separation) runs through the C ABI.
Tolerances: complex64 storage -> 2e-4 on W / output after a few iterations, 1e-4 on the loss.
"""
import numpy as np
import pytest

from conftest import load_golden, rel
from oracle import idlma as o_idlma, synth

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('name', ['idlma_gauss_d2', 'idlma_gauss_d1'])
def test_idlma_golden(cuda_device, name):
    from audio_source_separation_b200.sss.idlma import GaussIDLMA
    meta, i, o = load_golden(name)
    model = GaussIDLMA(domain=meta['domain'], normalize='projection-back')
    out = model(i['X'], iteration=meta['iteration'], dnn=synth.toy_dnn())
    assert rel(out, o['output']) < 2e-4
    assert rel(model.demix_filter, o['demix_filter']) < 2e-4
    assert rel(model.dnn_output, o['dnn_output']) < 2e-4
    assert len(model.loss) == meta['iteration'] + 1
    assert np.max(np.abs(np.array(model.loss) - o['loss']) / np.abs(o['loss'])) < 1e-4
    assert model.estimation is out
    assert model._handle.launch_count() > 0


@pytest.mark.parametrize('name', ['idlma_gauss_d2', 'idlma_gauss_d1'])
def test_idlma_space_model_alone(cuda_device, name):
    """update_space_model from injected variances (zeros exercise the eps floor), then the loss."""
    from audio_source_separation_b200.sss.idlma import GaussIDLMA
    meta, i, o = load_golden(name)
    model = GaussIDLMA(domain=meta['domain'], normalize='projection-back')
    model.input = i['X']
    model._reset(dnn=None)
    assert np.array_equal(model.demix_filter, np.tile(np.eye(i['X'].shape[0]), (i['X'].shape[1], 1, 1)))
    assert rel(model.estimation, i['X']) < 1e-6
    model.dnn_output = i['R0'].copy()
    model.update_space_model()
    assert rel(model.demix_filter, o['space_W1']) < 1e-4
    loss = model.compute_negative_loglikelihood()
    assert abs(loss - o['space_loss1']) < 1e-4 * abs(o['space_loss1'])


def test_idlma_against_oracle_larger(cuda_device):
    """A shape the fixtures do not cover (4 channels, ragged frame count), numpy callable instead of a torch module."""
    from audio_source_separation_b200.sss.idlma import GaussIDLMA
    C, F, T = 4, 65, 131
    X = synth.mix2(C, F, T, seed=6)

    def dnn(a):   # moving average over 3 bins, float64 throughout
        p = np.pad(a, ((0, 0), (1, 1), (0, 0)), mode='edge')
        return (p[:, :-2] + p[:, 1:-1] + p[:, 2:]) / 3 + 1e-4

    seen = []
    model = GaussIDLMA(domain=2, normalize='projection-back', reference_id=1, callback=lambda m: seen.append(m.demix_filter.copy()))
    out = model(X, iteration=4, dnn=dnn)
    want, st, loss = o_idlma.run(X, dnn, iteration=4, domain=2, reference_id=1)
    assert rel(out, want) < 2e-4 and rel(model.demix_filter, st['W']) < 2e-4
    assert np.max(np.abs(np.array(model.loss) - np.array(loss)) / np.abs(np.array(loss))) < 1e-4
    assert len(seen) == 4 and rel(seen[-1], st['W']) < 2e-4


def test_idlma_reference_quirks(cuda_device):
    """The default normalize='power' raises after the sweep (src/sss/idlma.py:159-160), normalize=False too (:161-162);
    presets of the filter are ignored by _reset (:34-36); the loss list keeps growing across calls (:15)."""
    from audio_source_separation_b200.sss.idlma import GaussIDLMA
    X = synth.mix2(2, 9, 24, seed=1)
    with pytest.raises(ValueError, match="Not support normalization based on power"):
        GaussIDLMA()(X, iteration=1, dnn=synth.toy_dnn())
    with pytest.raises(ValueError, match="Set normalize=True"):
        GaussIDLMA(normalize=False)(X, iteration=1, dnn=synth.toy_dnn())
    with pytest.raises(AssertionError):
        GaussIDLMA(domain=3)
    model = GaussIDLMA(normalize='projection-back')
    model(X, iteration=1, dnn=synth.toy_dnn(), demix_filter=synth.random_demix(2, 9, seed=2))
    first = model.demix_filter.copy()
    model(X, iteration=1, dnn=synth.toy_dnn())
    assert rel(model.demix_filter, first) < 1e-6
    assert len(model.loss) == 4
