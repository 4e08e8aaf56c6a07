"""GPU parity tests: the CUDA path (through the C ABI / the drop-in classes) against the golden fixtures
generated from the unmodified reference and against the oracle on seeded inputs.

Tolerances (the CUDA path stores complex64 / float32 and solves per-bin systems
in float64, the reference is complex128 throughout):
    single kernel, identical inputs ........ 1e-5 relative Frobenius (covariance), 1e-4 (IP rows)
    a few update_once from identical state . 2e-4 on W, Y, T, V;  loss 1e-4 relative
    100-iteration trajectories ............. 1e-3 on the projection-backed output, loss 1e-4 per iteration
    index / gate decisions ................. bit exact
"""
import numpy as np
import pytest

from conftest import load_golden, rel
from oracle import core, ilrma as o_ilrma, auxiva as o_auxiva, synth

pytestmark = pytest.mark.gpu

TOL_STATE = 2e-4
TOL_LOSS = 1e-4


def _loss_close(got, want, tol=TOL_LOSS):
    got, want = np.asarray(got), np.asarray(want)
    assert got.shape == want.shape
    scale = np.maximum(np.abs(want), 1e-3 * np.max(np.abs(want)))
    assert np.max(np.abs(got - want) / scale) < tol, (got, want)


# ------------------------------------------------------------------------------------------- primitives

def test_weighted_covariance_golden(cuda_device):
    from audio_source_separation_b200 import _lib
    meta, i, o = load_golden('primitives')
    U = _lib.weighted_covariance(i['X'], i['R'])
    assert rel(U, o['U']) < 1e-5
    assert np.max(np.abs(U - np.conj(np.swapaxes(U, -1, -2)))) == 0.0   # exactly Hermitian
    per_bin = np.linalg.norm((U - o['U']).reshape(U.shape[0], U.shape[1], -1), axis=-1) / np.linalg.norm(
        o['U'].reshape(U.shape[0], U.shape[1], -1), axis=-1)
    assert per_bin.max() < 1e-4


@pytest.mark.parametrize('C,F,T', [(2, 65, 130), (3, 33, 77), (4, 129, 512), (4, 17, 1000), (5, 9, 64), (8, 9, 300)])
def test_weighted_covariance_shapes(cuda_device, C, F, T):
    """Whole-tile and slab paths, odd frame counts (zero padded frame), every channel count family."""
    from audio_source_separation_b200 import _lib
    X = synth.mix2(C, F, T, seed=C + T)
    rng = np.random.default_rng(T)
    R = 10 ** rng.uniform(-6, 2, size=(C, F, T))
    U = _lib.weighted_covariance(X, R)
    want = core.weighted_covariance(X, R)
    assert rel(U, want) < 1e-5


def test_ip_update_matches_oracle(cuda_device):
    from audio_source_separation_b200 import _lib
    for C in (2, 3, 4, 6):
        F, T = 37, 60
        X = synth.mix2(C, F, T, seed=C)
        rng = np.random.default_rng(2)
        R = 10 ** rng.uniform(-3, 1, size=(C, F, T))
        U = core.weighted_covariance(X, R)
        W0 = synth.random_demix(C, F, seed=5)
        for floor in (False, True):
            Wo = W0.copy()
            gate_o = core.ip_rows(Wo, U, den_floor=1e-12 if floor else None)
            Wg, gate_g = _lib.ip_update(W0, U, floor_den=floor)
            assert np.array_equal(gate_g, gate_o)      # bit-exact gate decisions
            assert rel(Wg, Wo) < 1e-9                  # fp64 in-kernel


def test_ip_update_gate_near_threshold(cuda_device):
    from audio_source_separation_b200 import _lib
    rng = np.random.default_rng(3)
    C, F = 3, 24
    U = np.empty((C, F, C, C), dtype=np.complex128)
    conds = 10 ** np.linspace(9, 15, F)
    for n in range(C):
        for f in range(F):
            Q, _ = np.linalg.qr(rng.standard_normal((C, C)) + 1j * rng.standard_normal((C, C)))
            U[n, f] = (Q * np.array([1.0, 0.1, 1.0 / conds[f]])) @ Q.conj().T
    W0 = np.tile(np.eye(C, dtype=np.complex128), (F, 1, 1))
    Wo = W0.copy()
    gate_o = core.ip_rows(Wo, U)
    Wg, gate_g = _lib.ip_update(W0, U)
    assert np.array_equal(gate_g[0], gate_o[0])
    assert gate_g[0].any() and not gate_g[0].all()
    assert np.array_equal(Wg[~gate_o[0], 0], W0[~gate_o[0], 0])   # gated rows are kept bit for bit


def test_ip_update_singular_bin_raises(cuda_device):
    """An all-zero bin makes the reference raise LinAlgError."""
    from audio_source_separation_b200 import _lib
    C, F = 2, 5
    U = np.tile(np.eye(C, dtype=np.complex128), (C, F, 1, 1))
    U[:, 2] = 0
    W0 = np.tile(np.eye(C, dtype=np.complex128), (F, 1, 1))
    with pytest.raises(np.linalg.LinAlgError):
        _lib.ip_update(W0, U)
    with pytest.raises(np.linalg.LinAlgError):
        core.ip_rows(W0.copy(), U)


def test_projection_back_and_demix_golden(cuda_device):
    from audio_source_separation_b200 import _lib
    meta, i, o = load_golden('primitives')
    scale = _lib.projection_back_scale(i['X'], i['W'], 0)
    assert rel(scale, o['scale2']) < 1e-5
    for c in range(i['X'].shape[0]):
        assert rel(_lib.projection_back_scale(i['X'], i['W'], c), o['scale3'][c]) < 1e-5
    Y = _lib.demix(i['X'], i['W'])
    assert rel(Y, core.demix(i['X'], i['W'])) < 1e-6


def test_parallel_sort_and_pair_schedule_bit_exact(cuda_device):
    from audio_source_separation_b200.utils.utils_linalg import parallel_sort
    from audio_source_separation_b200.bss.ilrma import GaussILRMA
    meta, i, o = load_golden('primitives')
    assert np.array_equal(parallel_sort(i['eigvec'].swapaxes(-2, -1), order=o['order'], axis=-2), o['sorted'])
    for n_src in (2, 3, 4):
        m = GaussILRMA(n_basis=2, algorithm_spatial='IP2')
        m.n_sources = n_src
        seq = []
        for _ in range(7):
            m._select_update_pair()
            seq.append(m.update_pair)
        assert np.array_equal(np.array(seq), o['pairs{}'.format(n_src)])


# ------------------------------------------------------------------------------------------- ILRMA

ILRMA_CASES = ['ilrma_ip_power_d2', 'ilrma_ip_power_d1', 'ilrma_ip_pb_d2', 'ilrma_iss_power_d2', 'ilrma_iss_pb_d1',
               'ilrma_ip2_power_d2', 'ilrma_ip2_power_c2']


def _phase_align(W, ref):
    """IP2 rows are defined up to a per-row phase (eigenvector normalisation, SURVEY.md hard part 5)."""
    ph = np.sum(ref * np.conj(W), axis=-1, keepdims=True)
    ph = ph / np.maximum(np.abs(ph), 1e-300)
    return W * ph


@pytest.mark.parametrize('name', ILRMA_CASES)
def test_gauss_ilrma_golden(cuda_device, name):
    import warnings
    from audio_source_separation_b200.bss.ilrma import GaussILRMA
    meta, i, o = load_golden(name)
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        model = GaussILRMA(n_basis=meta['n_basis'], domain=meta['domain'], normalize=meta['normalize'],
                           algorithm_spatial=meta['algorithm_spatial'])
    out = model(i['X'], iteration=meta['iteration'], demix_filter=i['W0'], basis=i['T0'], activation=i['V0'])
    assert out.shape == o['output'].shape and out.dtype == np.complex128
    assert rel(out, o['output']) < TOL_STATE
    assert rel(model.basis, o['basis']) < TOL_STATE
    assert rel(model.activation, o['activation']) < TOL_STATE
    W = model.demix_filter
    if meta['algorithm_spatial'] in ('IP2', 'pairwise'):
        W = _phase_align(W, o['demix_filter'])
    assert rel(W, o['demix_filter']) < TOL_STATE
    _loss_close(model.loss, o['loss'])
    assert model.estimation is out


@pytest.mark.parametrize('spatial,K,C', [('IP', 2, 3), ('IP', 5, 4), ('ISS', 3, 2)])
def test_gauss_ilrma_partitioned(cuda_device, spatial, K, C):
    """Shared basis with latent source assignment (src/bss/ilrma.py:368-408, :313-320): golden fixture plus
    oracle runs for ISS and for n_basis above / below the kernels' accumulation chunk."""
    import warnings
    from audio_source_separation_b200.bss.ilrma import GaussILRMA
    meta, i, o = load_golden('ilrma_ip_power_part')
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        model = GaussILRMA(n_basis=meta['n_basis'], partitioning=True)
        out = model(i['X'], iteration=meta['iteration'], demix_filter=i['W0'], basis=i['T0'], activation=i['V0'], latent=i['Z0'])
        assert rel(out, o['output']) < TOL_STATE
        assert rel(model.basis, o['basis']) < TOL_STATE and rel(model.activation, o['activation']) < TOL_STATE
        assert rel(model.demix_filter, o['demix_filter']) < TOL_STATE
        _loss_close(model.loss, o['loss'])
        F, T = 33, 61
        X = synth.mix2(C, F, T, seed=C)
        rng = np.random.default_rng(3)
        W0 = np.tile(np.eye(C, dtype=np.complex128), (F, 1, 1))
        T0 = rng.random((F, K)).astype(np.float32).astype(np.float64)
        V0 = rng.random((K, T)).astype(np.float32).astype(np.float64)
        Z0 = rng.random((C, K)).astype(np.float32).astype(np.float64)
        Z0 = (Z0 / Z0.sum(axis=0)).astype(np.float32).astype(np.float64)
        model = GaussILRMA(n_basis=K, partitioning=True, algorithm_spatial=spatial)
        out = model(X, iteration=3, demix_filter=W0, basis=T0, activation=V0, latent=Z0)
    want, st, loss = o_ilrma.run(X, iteration=3, n_basis=K, spatial=spatial, partitioning=True, W=W0, T=T0, V=V0, Z=Z0)
    assert rel(out, want) < TOL_STATE
    assert rel(model.basis, st['T']) < TOL_STATE and rel(model.activation, st['V']) < TOL_STATE
    assert rel(model.latent, st['Z']) < TOL_STATE
    _loss_close(model.loss, loss)
    with pytest.raises(NotImplementedError):
        GaussILRMA(n_basis=2, partitioning=True, normalize='projection-back', recordable_loss=False)(X, iteration=1)


def test_gauss_ilrma_update_once_by_hand(cuda_device):
    """Users may drive update_once themselves after assigning input/state."""
    from audio_source_separation_b200.bss.ilrma import GaussILRMA
    meta, i, o = load_golden('ilrma_ip_power_d2')
    model = GaussILRMA(n_basis=meta['n_basis'], recordable_loss=False)
    model.input = i['X']
    model._reset(demix_filter=i['W0'], basis=i['T0'], activation=i['V0'])
    st = o_ilrma.init_state(i['X'], meta['n_basis'], W=i['W0'], T=i['T0'], V=i['V0'])
    for _ in range(2):
        model.update_once()
        o_ilrma.update_once(st)
        assert rel(model.demix_filter, st['W']) < TOL_STATE
        assert rel(model.estimation, st['Y']) < TOL_STATE
        assert rel(model.basis, st['T']) < TOL_STATE
        assert rel(model.activation, st['V']) < TOL_STATE
    # host assignment between iterations is honoured
    T_new = np.asarray(model.basis) * 1.5
    model.basis = T_new
    st['T'] = T_new.copy()
    model.update_once()
    o_ilrma.update_once(st)
    assert rel(model.basis, st['T']) < TOL_STATE and rel(model.demix_filter, st['W']) < TOL_STATE


def test_gauss_ilrma_seeded_dropin(cuda_device):
    """No injected state: the RNG must be consumed exactly like the reference (src/bss/ilrma.py:97-104)."""
    from audio_source_separation_b200.bss.ilrma import GaussILRMA
    meta, i, o = load_golden('ilrma_seeded_dropin')
    np.random.seed(meta['seed'])
    model = GaussILRMA(n_basis=meta['n_basis'])
    out = model(i['X'], iteration=meta['iteration'])
    assert rel(out, o['output']) < TOL_STATE
    assert rel(model.basis, o['basis']) < TOL_STATE
    _loss_close(model.loss, o['loss'])


def test_gauss_ilrma_callbacks_and_device_loop_agree(cuda_device):
    """The device-side loop (no loss, no callbacks) and the Python loop give the same result; callbacks see
    NumPy state after every iteration."""
    from audio_source_separation_b200.bss.ilrma import GaussILRMA
    meta, i, o = load_golden('ilrma_ip_power_d2')
    seen = []

    def cb(m):
        seen.append((m.demix_filter.shape, m.estimation.shape, m.basis.shape, float(np.abs(m.activation).sum())))

    a = GaussILRMA(n_basis=meta['n_basis'], callbacks=cb, recordable_loss=True)
    out_a = a(i['X'], iteration=meta['iteration'], demix_filter=i['W0'], basis=i['T0'], activation=i['V0'])
    b = GaussILRMA(n_basis=meta['n_basis'], recordable_loss=False)
    out_b = b(i['X'], iteration=meta['iteration'], demix_filter=i['W0'], basis=i['T0'], activation=i['V0'])
    assert len(seen) == meta['iteration'] + 1
    assert np.array_equal(out_a, out_b)
    assert b.loss is None and len(a.loss) == meta['iteration'] + 1


def test_tilrma_golden(cuda_device):
    from audio_source_separation_b200.bss.ilrma import tILRMA
    meta, i, o = load_golden('tilrma_nu5')
    model = tILRMA(n_basis=meta['n_basis'], nu=meta['nu'])
    out = model(i['X'], iteration=meta['iteration'], demix_filter=i['W0'], basis=i['T0'], activation=i['V0'])
    assert rel(out, o['output']) < TOL_STATE
    assert rel(model.basis, o['basis']) < TOL_STATE
    assert rel(model.activation, o['activation']) < TOL_STATE
    assert rel(model.demix_filter, o['demix_filter']) < TOL_STATE
    _loss_close(model.loss, o['loss'])


def test_gauss_ilrma_trajectory_small(cuda_device):
    """100 iterations on mix2(4,513,128): final output and the whole loss curve."""
    from audio_source_separation_b200.bss.ilrma import GaussILRMA
    X = synth.mix2(4, 513, 128, seed=0)
    W0, T0, V0 = synth.initial_state(4, 513, 128, 2, seed=7)
    model = GaussILRMA(n_basis=2)
    out = model(X, iteration=100, demix_filter=W0, basis=T0, activation=V0)
    want, st, loss = o_ilrma.run(X, iteration=100, n_basis=2, W=W0, T=T0, V=V0)
    assert rel(out, want) < 1e-3
    _loss_close(model.loss, loss)
    # known-answer values of the reference (Appendix D; the float32-rounded initial state moves them by < 1e-6)
    assert abs(model.loss[0] / 7.2776542359e5 - 1) < 1e-5
    assert abs(model.loss[100] / -1.9023653392e4 - 1) < 1e-3


def test_gauss_ilrma_headline_shape_one_update(cuda_device):
    """cfg3 (4 ch, 2049 bins, 512 frames, K = 2): one update_once and the loss against the oracle."""
    from audio_source_separation_b200.bss.ilrma import GaussILRMA
    X = synth.mix2(4, 2049, 512, seed=0)
    W0, T0, V0 = synth.initial_state(4, 2049, 512, 2, seed=7)
    model = GaussILRMA(n_basis=2, recordable_loss=False)
    model.input = X
    model._reset(demix_filter=W0, basis=T0, activation=V0)
    loss0 = model.compute_negative_loglikelihood()
    model.update_once()
    st = o_ilrma.init_state(X, 2, W=W0, T=T0, V=V0)
    want0 = o_ilrma.negative_loglikelihood(st)
    o_ilrma.update_once(st)
    assert abs(loss0 / want0 - 1) < 1e-5
    assert abs(loss0 / 1.0852043523e7 - 1) < 1e-5      # reference known answer (Appendix D)
    assert rel(model.demix_filter, st['W']) < TOL_STATE
    assert rel(model.basis, st['T']) < TOL_STATE
    assert rel(model.activation, st['V']) < TOL_STATE
    gate = model._handle.get_state(9, (4, 2049), np.int32)
    assert gate.all()
    loss1 = model.compute_negative_loglikelihood()
    assert abs(loss1 / o_ilrma.negative_loglikelihood(st) - 1) < 1e-4


def test_headline_shape_invariants(cuda_device):
    """Size-independent properties at the full headline size: the update is equivariant to a global rescaling
    of the mixture (power normalisation removes it) and the loss never increases."""
    from audio_source_separation_b200.bss.ilrma import GaussILRMA
    X = synth.mix2(4, 2049, 512, seed=1)
    W0, T0, V0 = synth.initial_state(4, 2049, 512, 2, seed=7)
    m1 = GaussILRMA(n_basis=2)
    out1 = m1(X, iteration=5, demix_filter=W0, basis=T0, activation=V0)
    assert all(b <= a + 1e-6 * abs(a) for a, b in zip(m1.loss[1:], m1.loss[2:]))
    m2 = GaussILRMA(n_basis=2, recordable_loss=False)
    out2 = m2(4.0 * X, iteration=5, demix_filter=W0, basis=T0, activation=V0)
    assert rel(out2, 4.0 * out1) < 1e-4   # projection back restores the input scale


# ------------------------------------------------------------------------------------------- AuxIVA

AUXIVA_CASES = ['auxiva_laplace_ip', 'auxiva_laplace_ip_c4', 'auxiva_gauss_ip', 'auxiva_laplace_iss', 'auxiva_gauss_iss',
                'auxiva_laplace_ip2']


@pytest.mark.parametrize('name', AUXIVA_CASES)
def test_auxiva_golden(cuda_device, name):
    from audio_source_separation_b200.bss.iva import AuxLaplaceIVA, AuxGaussIVA
    meta, i, o = load_golden(name)
    cls = AuxLaplaceIVA if meta['model'] == 'AuxLaplaceIVA' else AuxGaussIVA
    model = cls(algorithm_spatial=meta['algorithm_spatial'])
    out = model(i['X'], iteration=meta['iteration'], demix_filter=i['W0'])
    assert rel(out, o['output']) < TOL_STATE
    W = model.demix_filter
    if meta['algorithm_spatial'] in ('IP2', 'pairwise'):
        W = _phase_align(W, o['demix_filter'])
    assert rel(W, o['demix_filter']) < TOL_STATE
    _loss_close(model.loss, o['loss'])


def test_auxiva_cfg2_trajectory(cuda_device):
    """cfg2: AuxLaplaceIVA-IP, 2 ch, 1025 bins, 256 frames, 30 iterations (known answers: Appendix D)."""
    from audio_source_separation_b200.bss.iva import AuxLaplaceIVA
    X = synth.mix2(2, 1025, 256, seed=0)
    model = AuxLaplaceIVA()
    out = model(X, iteration=30)
    want, st, loss = o_auxiva.run(X, iteration=30, kind='laplace')
    assert rel(out, want) < 1e-3
    _loss_close(model.loss, loss)
    for idx, val in ((0, 2.5476010173e4), (1, -2.0310271328e6), (2, -2.7585783598e6), (30, -3.1106368346e6)):
        assert abs(model.loss[idx] / val - 1) < 1e-4
    assert abs(np.abs(out).sum() / 2.1658021777e5 - 1) < 1e-4


def test_auxgauss_ip2_not_implemented(cuda_device):
    from audio_source_separation_b200.bss.iva import AuxGaussIVA
    X = synth.mix2(2, 9, 20, seed=0)
    with pytest.raises(NotImplementedError):
        AuxGaussIVA(algorithm_spatial='IP2', recordable_loss=False)(X, iteration=1)


# ------------------------------------------------------------------------------------------- error paths

def test_bad_arguments_raise_before_gpu_work(cuda_device):
    from audio_source_separation_b200.bss.ilrma import GaussILRMA
    from audio_source_separation_b200.bss.iva import AuxLaplaceIVA
    with pytest.raises(AssertionError):
        GaussILRMA(domain=3)
    with pytest.raises(AssertionError):
        GaussILRMA(algorithm_spatial='IPA')
    with pytest.raises(ValueError):
        AuxLaplaceIVA(algorithm_spatial='nope')
    X = synth.mix2(2, 9, 20, seed=0)
    with pytest.raises(ValueError):
        GaussILRMA(n_basis=2, normalize='other', recordable_loss=False)(X, iteration=1)
    with pytest.raises(AssertionError):
        GaussILRMA().update_once()


def test_all_zero_bin_raises_linalgerror(cuda_device):
    from audio_source_separation_b200.bss.iva import AuxLaplaceIVA
    X = synth.mix2(2, 9, 20, seed=0)
    X[:, 4, :] = 0
    with pytest.raises(np.linalg.LinAlgError):
        AuxLaplaceIVA(recordable_loss=False)(X, iteration=2)


def test_update_spatial_model_ip_with_external_variances(cuda_device):
    """The spatial update as a stand-alone operator (what GaussIDLMA.update_space_model, src/sss/idlma.py:175-210, runs with
    DNN variances): covariance + IP sweep + demixing from the stateless C entry points."""
    from audio_source_separation_b200.bss.ilrma import update_spatial_model_ip
    C, F, T = 3, 21, 70
    X = synth.mix2(C, F, T, seed=9)
    rng = np.random.default_rng(4)
    R = 10 ** rng.uniform(-3, 1, size=(C, F, T))
    R[0, 0, :5] = 0.0                                   # exercises the eps floor
    W0 = synth.random_demix(C, F, seed=2)
    W, Y = update_spatial_model_ip(X, W0, R)
    Rf = np.maximum(R, 1e-12)
    Wo = W0.copy()
    core.ip_rows(Wo, core.weighted_covariance(X, Rf))
    assert rel(W, Wo) < 1e-4
    assert rel(Y, core.demix(X, Wo)) < 1e-4
