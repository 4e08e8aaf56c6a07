"""CPU: host-side logic of the AuxLaplaceIVA / AuxGaussIVA drop-ins with the device answered by the oracle
(tests/fake_handle.py): the six reference fixtures through the device loop and through the callback loop, the IP2 pair
schedule, the ISS filter bookkeeping and `apply_projection_back`."""
import numpy as np
import pytest

from conftest import load_golden, rel

CASES = ['auxiva_laplace_ip', 'auxiva_laplace_ip_c4', 'auxiva_gauss_ip', 'auxiva_laplace_iss', 'auxiva_gauss_iss',
         'auxiva_laplace_ip2']


@pytest.fixture
def iva(monkeypatch):
    from audio_source_separation_b200 import _lib
    from fake_handle import FakeAuxIVAHandle
    FakeAuxIVAHandle.instances.clear()
    monkeypatch.setattr(_lib, 'Handle', FakeAuxIVAHandle)
    from audio_source_separation_b200.bss import iva as mod
    return mod, FakeAuxIVAHandle


def _cls(mod, meta):
    return mod.AuxLaplaceIVA if meta['model'] == 'AuxLaplaceIVA' else mod.AuxGaussIVA


@pytest.mark.parametrize('name', CASES)
@pytest.mark.parametrize('with_callback', [False, True])
def test_reproduces_the_reference(iva, name, with_callback):
    mod, fake = iva
    meta, i, o = load_golden(name)
    seen = []
    cb = (lambda m: seen.append((m.demix_filter.shape, len(m.loss)))) if with_callback else None
    model = _cls(mod, meta)(algorithm_spatial=meta['algorithm_spatial'], callbacks=cb)
    out = model(i['X'], iteration=meta['iteration'], demix_filter=i['W0'])
    tol = 1e-6 if meta['algorithm_spatial'] in ('IP2', 'pairwise') else 1e-8
    assert rel(out, o['output']) < tol and rel(model.loss, o['loss']) < tol
    assert rel(model.demix_filter, o['demix_filter']) < tol
    h = fake.instances[-1]
    assert h.calls.count('set_input') == 1
    C, F, _ = i['X'].shape
    if with_callback:
        assert seen == [((F, C, C), k + 1) for k in range(meta['iteration'] + 1)]
        assert h.calls.count('update_once') == meta['iteration'] and ('run_record', meta['iteration']) not in h.calls
    else:
        assert ('run_record', meta['iteration']) in h.calls


def test_without_projection_back_and_loss(iva):
    mod, fake = iva
    from oracle import auxiva as o_auxiva
    meta, i, o = load_golden('auxiva_laplace_ip')
    model = mod.AuxLaplaceIVA(apply_projection_back=False, recordable_loss=False)
    out = model(i['X'], iteration=2)
    want, _, _ = o_auxiva.run(i['X'], iteration=2, kind='laplace', apply_projection_back=False, record_loss=False)
    assert rel(out, want) < 1e-10 and model.loss is None
    assert ('run', 2) in fake.instances[-1].calls


def test_unsupported_combinations(iva):
    mod, fake = iva
    meta, i, o = load_golden('auxiva_gauss_ip')
    with pytest.raises(NotImplementedError):
        mod.AuxGaussIVA(algorithm_spatial='IP2')(i['X'], iteration=1)        # src/bss/iva.py:777-778
    with pytest.raises(ValueError):
        mod.AuxLaplaceIVA(algorithm_spatial='nonsense')
