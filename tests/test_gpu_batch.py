"""Batches of independent mixtures (our extension; the reference has no batch axis): a batched run must equal
independent single-mixture runs, through the resident-batch API and through the pipelined whole-job call."""
import numpy as np
import pytest

from conftest import rel
from oracle import ilrma as o_ilrma, synth

pytestmark = pytest.mark.gpu


def _inputs(B, C, F, T, K):
    X = np.stack([synth.mix2(C, F, T, seed=20 + b) for b in range(B)])
    rng = np.random.default_rng(3)
    T0 = rng.random((B, C, F, K)).astype(np.float32).astype(np.float64)
    V0 = rng.random((B, C, K, T)).astype(np.float32).astype(np.float64)
    return X, T0, V0


@pytest.mark.parametrize('C,F,T,K', [(4, 65, 200, 2), (2, 33, 131, 3)])
def test_batch_equals_loop(cuda_device, C, F, T, K):
    from audio_source_separation_b200.batch import BatchedGaussILRMA
    from audio_source_separation_b200.bss.ilrma import GaussILRMA
    B, n_iter = 5, 4
    X, T0, V0 = _inputs(B, C, F, T, K)
    W0 = np.tile(np.eye(C, dtype=np.complex128), (F, 1, 1))
    singles = []
    for b in range(B):
        m = GaussILRMA(n_basis=K, recordable_loss=False)
        singles.append(m(X[b], iteration=n_iter, demix_filter=W0, basis=T0[b], activation=V0[b]))
    singles = np.stack(singles)
    batched = BatchedGaussILRMA(n_basis=K)
    out = batched(X, iteration=n_iter, basis=T0, activation=V0)
    assert out.shape == singles.shape
    assert rel(out, singles) < 1e-5
    assert batched.compute_negative_loglikelihood().shape == (B,)
    # pipelined whole-job call: sub-batches on their own streams, complex64 output
    for pipeline in (1, 2, 5, [1, 3, 1], 'ramp'):
        out2 = BatchedGaussILRMA(n_basis=K).separate_batch(X.astype(np.complex64), iteration=n_iter, basis=T0, activation=V0,
                                                          pipeline=pipeline)
        assert out2.dtype == np.complex64 and rel(out2, singles) < 1e-5
    # and against the oracle for the first two mixtures
    for b in range(2):
        want, _, _ = o_ilrma.run(X[b], iteration=n_iter, n_basis=K, W=W0, T=T0[b], V=V0[b], record_loss=False)
        assert rel(out[b], want) < 1e-3


def test_batch_handles_are_independent(cuda_device):
    """Changing one mixture of the batch must not change the others (no cross-mixture reduction leaks)."""
    from audio_source_separation_b200.batch import BatchedGaussILRMA
    C, F, T, K, B = 3, 40, 96, 2, 4
    X, T0, V0 = _inputs(B, C, F, T, K)
    a = BatchedGaussILRMA(n_basis=K)(X, iteration=3, basis=T0, activation=V0)
    X2 = X.copy()
    X2[2] *= 3.0
    b = BatchedGaussILRMA(n_basis=K)(X2, iteration=3, basis=T0, activation=V0)
    for i in (0, 1, 3):
        assert np.array_equal(a[i], b[i])
    assert not np.array_equal(a[2], b[2])


@pytest.mark.parametrize('kind', ['ilrma_ip', 'ilrma_iss', 'ilrma_part', 'auxiva', 'fastmnmf', 'nmf'])
def test_graph_replay_equals_eager_loop(cuda_device, kind, monkeypatch):
    """bss_run replays long loops from a CUDA graph of two iterations; the result must be bit-identical to the eager
    loop (same kernels, same order), including an odd number of iterations."""
    import warnings
    from audio_source_separation_b200 import _lib
    rng = np.random.default_rng(0)
    n_iter = 25

    def build():
        C, F, T, K = 3, 40, 150, 2
        X = synth.mix2(C, F, T, seed=4)
        common = dict(n_batch=1, n_channels=C, n_sources=C, n_bins=F, n_frames=T, n_basis=K)
        if kind == 'nmf':
            h = _lib.Handle(method=_lib.NMF_EUC, n_batch=1, n_channels=1, n_sources=1, n_bins=F, n_frames=T, n_basis=K)
            h.set_state(_lib.STATE_TARGET, np.abs(X[0]) ** 2, np.float64)
            h.set_state(_lib.STATE_BASIS, np.random.default_rng(1).random((F, K)), np.float64)
            h.set_state(_lib.STATE_ACTIVATION, np.random.default_rng(2).random((K, T)), np.float64)
            return h, [(_lib.STATE_BASIS, (F, K), np.float64), (_lib.STATE_ACTIVATION, (K, T), np.float64)]
        if kind == 'fastmnmf':
            h = _lib.Handle(method=_lib.FAST_MNMF, **common)
            h.set_input(X)
            h.reset_spatial()
            h.set_state(_lib.STATE_BASIS, np.random.default_rng(1).random((C, F, K)), np.float64)
            h.set_state(_lib.STATE_ACTIVATION, np.random.default_rng(2).random((C, K, T)), np.float64)
            return h, [(_lib.STATE_DIAGONALIZER, (F, C, C), np.complex128), (_lib.STATE_SPATIAL, (C, F, C), np.float64),
                       (_lib.STATE_BASIS, (C, F, K), np.float64)]
        if kind == 'auxiva':
            h = _lib.Handle(method=_lib.AUX_LAPLACE_IVA, normalize=_lib.NORMALIZE_NONE, **common)
            h.set_input(X)
            h.reset_spatial()
            return h, [(_lib.STATE_DEMIX_FILTER, (F, C, C), np.complex128)]
        part = kind == 'ilrma_part'
        h = _lib.Handle(method=_lib.GAUSS_ILRMA, spatial=_lib.SPATIAL_ISS if kind == 'ilrma_iss' else _lib.SPATIAL_IP,
                        partitioning=1 if part else 0, **common)
        h.set_input(X)
        h.reset_spatial()
        if part:
            Z = np.random.default_rng(3).random((C, K)) + 0.5
            h.set_state(_lib.STATE_LATENT, Z / Z.sum(axis=0), np.float64)
            h.set_state(_lib.STATE_BASIS, np.random.default_rng(1).random((F, K)), np.float64)
            h.set_state(_lib.STATE_ACTIVATION, np.random.default_rng(2).random((K, T)), np.float64)
            return h, [(_lib.STATE_DEMIX_FILTER, (F, C, C), np.complex128), (_lib.STATE_BASIS, (F, K), np.float64),
                       (_lib.STATE_LATENT, (C, K), np.float64)]
        h.set_state(_lib.STATE_BASIS, np.random.default_rng(1).random((C, F, K)), np.float64)
        h.set_state(_lib.STATE_ACTIVATION, np.random.default_rng(2).random((C, K, T)), np.float64)
        return h, [(_lib.STATE_ESTIMATION, (C, F, T), np.complex128), (_lib.STATE_BASIS, (C, F, K), np.float64),
                   (_lib.STATE_ACTIVATION, (C, K, T), np.float64)]

    results = []
    for no_graph in (False, True):
        if no_graph:
            monkeypatch.setenv('BSSGPU_NO_GRAPH', '1')
        else:
            monkeypatch.delenv('BSSGPU_NO_GRAPH', raising=False)
        h, states = build()
        n0 = h.launch_count()
        h.run(n_iter)
        h.synchronize()
        results.append(([h.get_state(w, shape, dt) for w, shape, dt in states], h.launch_count() - n0))
        h.close()
    (graph_states, graph_launches), (eager_states, eager_launches) = results
    assert graph_launches == eager_launches          # replayed launches are counted
    for a, b in zip(graph_states, eager_states):
        assert np.array_equal(a, b)


@pytest.mark.parametrize('kind', ['ilrma_ip', 'ilrma_iss', 'auxiva'])
def test_cached_graph_serves_later_runs(cuda_device, kind, monkeypatch):
    """A handle that runs job after job keeps its executable graph (bss_run re-captures only when the device pointers the
    kernels were recorded with changed, e.g. after an odd number of iterations swapped the basis buffers): the 2nd, 3rd,
    4th job on one handle must be bit-identical to the same job run eagerly on a fresh handle."""
    from audio_source_separation_b200 import _lib
    C, F, T, K = 3, 40, 150, 2
    X = synth.mix2(C, F, T, seed=4)
    X2 = synth.mix2(C, F, T, seed=5)
    T0 = np.random.default_rng(1).random((C, F, K))
    V0 = np.random.default_rng(2).random((C, K, T))
    common = dict(n_batch=1, n_channels=C, n_sources=C, n_bins=F, n_frames=T, n_basis=K)

    def make():
        if kind == 'auxiva':
            return _lib.Handle(method=_lib.AUX_LAPLACE_IVA, normalize=_lib.NORMALIZE_NONE, **common)
        return _lib.Handle(method=_lib.GAUSS_ILRMA, spatial=_lib.SPATIAL_ISS if kind == 'ilrma_iss' else _lib.SPATIAL_IP, **common)

    def job(h, x, n_iter):
        h.set_input(x)
        h.reset_spatial()
        if kind != 'auxiva':
            h.set_state(_lib.STATE_BASIS, T0, np.float64)
            h.set_state(_lib.STATE_ACTIVATION, V0, np.float64)
        h.run(n_iter)
        return h.separate((C, F, T), np.complex128, projection_back=True)

    monkeypatch.delenv('BSSGPU_NO_GRAPH', raising=False)
    h = make()
    plan = [(X, 24), (X2, 24), (X, 25), (X2, 24), (X, 24)]   # the odd job leaves the basis buffers swapped
    got = [job(h, x, n) for x, n in plan]
    h.close()
    monkeypatch.setenv('BSSGPU_NO_GRAPH', '1')
    for (x, n), g in zip(plan, got):
        e = make()
        want = job(e, x, n)
        e.close()
        assert np.array_equal(g, want)
