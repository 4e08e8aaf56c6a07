"""CPU: host-side logic of the GaussIDLMA drop-in (src/sss/idlma.py:88-258) with the device answered by the oracle
(tests/fake_handle.py).  The GPU suite (tests/test_gpu_idlma.py) runs the same fixtures through the real C ABI."""
import numpy as np
import pytest

from conftest import load_golden, rel
from oracle import synth


@pytest.fixture
def idlma(monkeypatch):
    from audio_source_separation_b200 import _lib
    from fake_handle import FakeIDLMAHandle
    FakeIDLMAHandle.instances.clear()
    monkeypatch.setattr(_lib, 'Handle', FakeIDLMAHandle)
    from audio_source_separation_b200.sss import idlma as mod
    return mod, FakeIDLMAHandle


@pytest.mark.parametrize('name', ['idlma_gauss_d2', 'idlma_gauss_d1'])
def test_call_reproduces_the_reference(idlma, name):
    mod, fake = idlma
    meta, i, o = load_golden(name)
    seen = []
    model = mod.GaussIDLMA(domain=meta['domain'], normalize='projection-back', callback=lambda m: seen.append(len(m.loss)))
    out = model(i['X'], iteration=meta['iteration'], dnn=synth.toy_dnn())
    assert rel(out, o['output']) < 1e-6 and rel(model.demix_filter, o['demix_filter']) < 1e-6
    assert rel(model.dnn_output, o['dnn_output']) < 1e-6 and rel(model.loss, o['loss']) < 1e-6
    assert model.estimation is out and seen == list(range(2, meta['iteration'] + 2))
    h = fake.instances[-1]
    # the mixture is uploaded once; per iteration: variances, spatial sweep, normalisation
    assert h.calls.count('set_input') == 1
    assert h.calls.count('update_once') == meta['iteration'] == h.calls.count('normalize')
    from audio_source_separation_b200 import _lib
    assert h.calls.count(('set_state', _lib.STATE_VARIANCE)) == meta['iteration'] + 1   # + the all-ones start


@pytest.mark.parametrize('name', ['idlma_gauss_d2', 'idlma_gauss_d1'])
def test_space_model_alone_and_variance_upload(idlma, name):
    mod, fake = idlma
    from audio_source_separation_b200 import _lib
    meta, i, o = load_golden(name)
    model = mod.GaussIDLMA(domain=meta['domain'], normalize='projection-back')
    model.input = i['X']
    model._reset(dnn=None)
    assert model.dnn is None and np.array_equal(model.dnn_output, np.ones(i['X'].shape))
    model.dnn_output = i['R0'].copy()
    model.update_space_model()
    assert rel(model.demix_filter, o['space_W1']) < 1e-9
    loss = model.compute_negative_loglikelihood()
    assert abs(loss - o['space_loss1']) < 1e-6 * abs(o['space_loss1'])
    h = fake.instances[-1]
    n_up = h.calls.count(('set_state', _lib.STATE_VARIANCE))
    model.compute_negative_loglikelihood()                      # unchanged variances are not uploaded again
    assert h.calls.count(('set_state', _lib.STATE_VARIANCE)) == n_up
    model.dnn_output[1, 2, 3] *= 2.0                            # in-place edits are noticed
    model.compute_negative_loglikelihood()
    assert h.calls.count(('set_state', _lib.STATE_VARIANCE)) == n_up + 1
    model.update_once(is_source_model_update=False)             # keeps the injected variances
    assert h.calls[-1] == 'normalize'


def test_reference_quirks(idlma):
    mod, fake = idlma
    X = synth.mix2(2, 9, 24, seed=1)
    with pytest.raises(ValueError, match="Not support normalization based on power"):
        mod.GaussIDLMA()(X, iteration=1, dnn=synth.toy_dnn())
    assert 'update_once' in fake.instances[-1].calls            # the sweep ran before the reference raises (:159-160)
    with pytest.raises(ValueError, match="Set normalize=True"):
        mod.GaussIDLMA(normalize=False)(X, iteration=1, dnn=synth.toy_dnn())
    with pytest.raises(AssertionError):
        mod.GaussIDLMA(domain=0.5)
    with pytest.raises(NotImplementedError):
        mod.IDLMAbase().update_once()
    model = mod.GaussIDLMA(normalize='projection-back')
    model(X, iteration=1, dnn=synth.toy_dnn(), demix_filter=synth.random_demix(2, 9, seed=2))   # preset ignored (:34-36)
    first = model.demix_filter.copy()
    model(X, iteration=1, dnn=lambda a: synth.dnn_as_callable(synth.toy_dnn())(a))              # plain callable
    assert rel(model.demix_filter, first) < 1e-9 and len(model.loss) == 4
