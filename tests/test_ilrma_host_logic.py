"""CPU: host-side logic of the GaussILRMA drop-in (reset order, presets, loss recording with and without callbacks, IP2
pair schedule, ISS filter bookkeeping, the lazy attribute protocol of `_model.DeviceModel`) with the device answered by
the oracle (tests/fake_handle.py).  tests/test_gpu_parity.py runs the same fixtures through the real C ABI."""
import warnings

import numpy as np
import pytest

from conftest import load_golden, rel

CASES = ['ilrma_ip_power_d2', 'ilrma_ip_power_d1', 'ilrma_ip_pb_d2', 'ilrma_iss_power_d2', 'ilrma_iss_pb_d1',
         'ilrma_ip2_power_d2', 'ilrma_ip2_power_c2']


@pytest.fixture
def ilrma(monkeypatch):
    from audio_source_separation_b200 import _lib
    from fake_handle import FakeILRMAHandle
    FakeILRMAHandle.instances.clear()
    monkeypatch.setattr(_lib, 'Handle', FakeILRMAHandle)
    from audio_source_separation_b200.bss import ilrma as mod
    return mod, FakeILRMAHandle


def _model(mod, meta, **kw):
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        return mod.GaussILRMA(n_basis=meta['n_basis'], domain=meta['domain'], normalize=meta['normalize'],
                              algorithm_spatial=meta['algorithm_spatial'], **kw)


def _check(model, out, o, tol):
    assert rel(out, o['output']) < tol and rel(model.loss, o['loss']) < tol
    assert rel(model.basis, o['basis']) < tol and rel(model.activation, o['activation']) < tol
    assert rel(model.demix_filter, o['demix_filter']) < tol


@pytest.mark.parametrize('name', CASES)
def test_device_loop_reproduces_the_reference(ilrma, name):
    """No callbacks: the whole loop is handed to the device in one call (run_record)."""
    mod, fake = ilrma
    meta, i, o = load_golden(name)
    model = _model(mod, meta)
    out = model(i['X'], iteration=meta['iteration'], demix_filter=i['W0'], basis=i['T0'], activation=i['V0'])
    tol = 1e-6 if meta['algorithm_spatial'] in ('IP2', 'pairwise') else 1e-8
    _check(model, out, o, tol)
    h = fake.instances[-1]
    assert ('run_record', meta['iteration']) in h.calls and 'update_once' in h.calls
    assert h.calls.count('set_input') == 1
    if meta['algorithm_spatial'] in ('IP2', 'pairwise'):   # the host's schedule ends where the reference's does
        pair = None
        from oracle import core
        for _ in range(meta['iteration']):
            pair = core.next_update_pair(pair, i['X'].shape[0])
        assert model.update_pair == pair


@pytest.mark.parametrize('name', CASES)
def test_callback_loop_reproduces_the_reference(ilrma, name):
    """With a callback the host drives update_once itself and the callback sees NumPy state after every iteration."""
    mod, fake = ilrma
    meta, i, o = load_golden(name)
    seen = []
    model = _model(mod, meta, callbacks=lambda m: seen.append((m.demix_filter.shape, m.estimation.shape, len(m.loss))))
    out = model(i['X'], iteration=meta['iteration'], demix_filter=i['W0'], basis=i['T0'], activation=i['V0'])
    tol = 1e-6 if meta['algorithm_spatial'] in ('IP2', 'pairwise') else 1e-8
    _check(model, out, o, tol)
    C, F, T = i['X'].shape
    assert seen == [((F, C, C), (C, F, T), k + 1) for k in range(meta['iteration'] + 1)]


def test_seeded_dropin_draws_like_the_reference(ilrma):
    mod, fake = ilrma
    meta, i, o = load_golden('ilrma_seeded_dropin')
    np.random.seed(meta['seed'])
    model = mod.GaussILRMA(n_basis=meta['n_basis'])
    out = model(i['X'], iteration=meta['iteration'])
    assert rel(out, o['output']) < 1e-8 and rel(model.loss, o['loss']) < 1e-8


def test_attribute_protocol(ilrma):
    """State lives on the device: reads are fetched lazily and cached, assignments and in-place edits are uploaded before the
    next device operation, and presets are copied (src/bss/ilrma.py:67-104)."""
    mod, fake = ilrma
    from audio_source_separation_b200 import _lib
    meta, i, o = load_golden('ilrma_ip_power_d2')
    model = _model(mod, meta, recordable_loss=False)
    assert not hasattr(model, 'basis') and model.loss is None
    T0 = i['T0'].copy()
    model.input = i['X']
    model._reset(demix_filter=i['W0'], basis=T0, activation=i['V0'])
    T0 *= 0.0                                                     # the preset was copied
    h = fake.instances[-1]
    assert rel(h.T, i['T0']) == 0.0
    model.update_once()
    n_get = h.calls.count(('get_state', _lib.STATE_BASIS))
    b1 = model.basis
    b2 = model.basis                                               # cached until the device moves on
    assert b1 is b2 and h.calls.count(('get_state', _lib.STATE_BASIS)) == n_get + 1
    b1 *= 2.0                                                      # in-place edit of a fetched array ...
    model.update_once()                                            # ... is uploaded before the update
    want = fake(**h.cfg)
    want.set_input(i['X'])
    want.set_state(_lib.STATE_DEMIX_FILTER, i['W0'], np.complex128)
    want.set_state(_lib.STATE_BASIS, i['T0'], np.float64)
    want.set_state(_lib.STATE_ACTIVATION, i['V0'], np.float64)
    want.update_once()
    want.T = want.T * 2.0
    want.update_once()
    assert rel(model.basis, want.T) < 1e-12 and rel(model.demix_filter, want.W) < 1e-12
    model.activation = np.ones_like(want.V)                        # plain assignment
    model.update_once()
    want.V = np.ones_like(want.V)
    want.update_once()
    assert rel(model.activation, want.V) < 1e-12
    with pytest.raises(ValueError):
        model.basis = np.ones((2, 2))
        model.update_once()
