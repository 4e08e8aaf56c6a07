"""GPU parity of the STFT feed (src/transform/stft.py, i.e. scipy.signal.stft / istft) and of the classes built on it.
The kernels compute in float32: 2e-6 relative Frobenius error on the spectrogram and on the reconstruction."""
import numpy as np
import pytest
from scipy import signal as ss

from conftest import load_golden, rel

pytestmark = pytest.mark.gpu

TOL = 2e-6


def test_stft_istft_golden(cuda_device):
    from audio_source_separation_b200.transform.stft import stft, istft
    meta, i, o = load_golden('ilrma_consistent')
    n, h = meta['fft_size'], meta['hop_size']
    Z = stft(i['x'], fft_size=n, hop_size=h)
    assert Z.shape == o['stft'].shape and Z.dtype == np.complex128
    assert rel(Z, o['stft']) < TOL
    y = istft(o['stft'], fft_size=n, hop_size=h, length=meta['n_samples'])
    assert y.shape == o['istft'].shape
    assert rel(y, o['istft']) < TOL
    assert rel(y, i['x']) < TOL          # perfect reconstruction (NOLA holds for hann at 75 % overlap)


@pytest.mark.parametrize('n_samples,fft,hop,window', [(66, 8, 2, 'hann'), (1000, 64, 16, 'hamming'), (16000, 1024, 256, 'hann'),
                                                      (40001, 4096, 2048, 'hann'), (5000, 256, 100, 'hann'), (600, 512, 128, 'hann')])
def test_stft_matches_scipy(cuda_device, n_samples, fft, hop, window):
    """The reference's own settings ((4096, 2048) for BSS, (1024, 256) for NMF, the (8, 2) of its _test), a hop that does not
    divide the frame, a signal barely longer than one frame, odd lengths; batches with leading axes."""
    from audio_source_separation_b200.transform.stft import stft, istft
    rng = np.random.default_rng(n_samples)
    x = rng.standard_normal((2, 3, n_samples))
    Z = stft(x, fft_size=fft, hop_size=hop, window_fn=window)
    want = ss.stft(x, nperseg=fft, noverlap=fft - hop, window=window)[2]
    assert Z.shape == want.shape
    assert rel(Z, want) < TOL
    y = istft(want, fft_size=fft, hop_size=hop, window_fn=window)
    y_want = ss.istft(want, nperseg=fft, noverlap=fft - hop, window=window)[1]
    assert y.shape == y_want.shape
    assert rel(y, y_want) < TOL


def test_stft_rejects_non_power_of_two(cuda_device):
    from audio_source_separation_b200.transform.stft import stft
    with pytest.raises(NotImplementedError):
        stft(np.zeros(1000), fft_size=1000, hop_size=250)


def test_consistent_ilrma_golden(cuda_device):
    from audio_source_separation_b200.bss.ilrma import ConsistentGaussILRMA
    meta, i, o = load_golden('ilrma_consistent')
    model = ConsistentGaussILRMA(n_basis=meta['n_basis'], fft_size=meta['fft_size'], hop_size=meta['hop_size'])
    out = model(i['X'], iteration=meta['iteration'], demix_filter=i['W0'], basis=i['T0'], activation=i['V0'])
    assert rel(out, o['output']) < 2e-4
    assert rel(model.basis, o['basis']) < 2e-4 and rel(model.demix_filter, o['demix_filter']) < 2e-4
    assert np.max(np.abs(np.array(model.loss) - o['loss']) / np.abs(o['loss'])) < 1e-4
    with pytest.raises(ValueError):
        ConsistentGaussILRMA(n_basis=2)


def test_waveform_feed_equals_spectrogram_feed(cuda_device):
    """bss_set_input_waveform: the STFT goes straight into the handle's bin tiles; the update loop must see the same mixture
    as when the host computes the spectrogram (odd and even frame counts, multi-block tiles)."""
    from audio_source_separation_b200 import _lib
    for n_samples, fft, hop in ((4000, 64, 16), (4100, 128, 32)):
        rng = np.random.default_rng(fft)
        B, C, K = 2, 3, 2
        x = rng.standard_normal((B, C, n_samples))
        X = ss.stft(x, nperseg=fft, noverlap=fft - hop, window='hann')[2]        # (B,C,F,T)
        F, T = X.shape[2:]
        assert T > 128
        T0, V0 = rng.random((B, C, F, K)), rng.random((B, C, K, T))
        outs = []
        for feed in ('spectrogram', 'waveform'):
            h = _lib.Handle(method=_lib.GAUSS_ILRMA, n_batch=B, n_channels=C, n_sources=C, n_bins=F, n_frames=T, n_basis=K)
            h.reset_spatial()
            h.set_state(_lib.STATE_BASIS, T0, np.float64)
            h.set_state(_lib.STATE_ACTIVATION, V0, np.float64)
            if feed == 'spectrogram':
                h.set_input(X)
            else:
                h.set_input_waveform(x, fft, hop, ss.get_window('hann', fft))
            h.run(3)
            outs.append(h.separate((B, C, F, T), np.complex128, projection_back=True))
            h.close()
        assert rel(outs[1], outs[0]) < 1e-4


def test_waveform_in_waveform_out(cuda_device):
    """BatchedGaussILRMA.separate_waveforms: STFT, update loop and ISTFT on the device against the host pipeline
    scipy.stft -> oracle ILRMA -> scipy.istft."""
    from audio_source_separation_b200.batch import BatchedGaussILRMA
    from oracle import ilrma as o_ilrma
    rng = np.random.default_rng(3)
    B, C, K, n_samples, fft, hop = 2, 2, 2, 6000, 128, 32
    s = rng.standard_normal((B, C, n_samples)) * np.array([1.0, 0.3])[None, :, None]
    A = np.array([[1.0, 0.6], [0.4, 1.0]])
    x = np.einsum('ij,bjt->bit', A, s)
    X = ss.stft(x, nperseg=fft, noverlap=fft - hop, window='hann')[2]
    F, T = X.shape[2:]
    T0, V0 = rng.random((B, C, F, K)), rng.random((B, C, K, T))
    y = BatchedGaussILRMA(n_basis=K).separate_waveforms(x, fft, hop, iteration=5, basis=T0, activation=V0)
    W0 = np.tile(np.eye(C, dtype=np.complex128), (F, 1, 1))
    for b in range(B):
        Xb = X[b].astype(np.complex64).astype(np.complex128)
        Y, _, _ = o_ilrma.run(Xb, iteration=5, n_basis=K, W=W0, T=T0[b].astype(np.float32).astype(np.float64),
                              V=V0[b].astype(np.float32).astype(np.float64), record_loss=False)
        want = ss.istft(Y, nperseg=fft, noverlap=fft - hop, window='hann')[1]
        assert y[b].shape == want.shape
        assert rel(y[b], want) < 1e-3


def test_pipelined_waveform_job_equals_the_single_handle_call(cuda_device):
    """BatchedGaussILRMA.separate_waveform_batch (sub-batches on their own handles / streams / host threads, float32 and
    float64 waveforms, scratch kept on the handles from job to job) against separate_waveforms on one handle."""
    from audio_source_separation_b200.batch import BatchedGaussILRMA
    rng = np.random.default_rng(5)
    B, C, K, n_samples, fft, hop = 6, 3, 2, 5000, 256, 128
    x = rng.standard_normal((B, C, n_samples))
    F = fft // 2 + 1
    T = len(ss.stft(x[0, 0], nperseg=fft, noverlap=fft - hop)[1])
    T0, V0 = rng.random((B, C, F, K)), rng.random((B, C, K, T))
    want = BatchedGaussILRMA(n_basis=K).separate_waveforms(x, fft, hop, iteration=12, basis=T0, activation=V0)
    model = BatchedGaussILRMA(n_basis=K)
    for pipeline in (1, 3, [1, 4, 1]):
        for job in range(2):   # the second job reuses handles, cached FFT tables, scratch buffers and the captured graph
            got = model.separate_waveform_batch(x, fft, hop, iteration=12, basis=T0, activation=V0, pipeline=pipeline)
            assert got.shape == want.shape and got.dtype == np.float64
            assert rel(got, want) < 1e-5
    got32 = model.separate_waveform_batch(x.astype(np.float32), fft, hop, iteration=12, basis=T0, activation=V0, pipeline=2)
    assert got32.dtype == np.float32 and rel(got32, want) < 1e-4
    with pytest.raises(ValueError):
        model.separate_waveform_batch(x, fft, hop, iteration=1, basis=T0[:, :, :-1], activation=V0)


def test_pcm16_waveform_feed(cuda_device):
    """int16 PCM in (scaled by 1 / 32768 on the device, as the reference's notebooks do after wavfile.read) equals the float
    feed of the same samples, through the single-handle and the pipelined job."""
    from audio_source_separation_b200.batch import BatchedGaussILRMA
    rng = np.random.default_rng(6)
    B, C, K, n_samples, fft, hop = 4, 2, 2, 4000, 256, 64
    pcm = rng.integers(-20000, 20000, size=(B, C, n_samples)).astype(np.int16)
    xf = pcm.astype(np.float32) / 32768
    F = fft // 2 + 1
    T = len(ss.stft(xf[0, 0], nperseg=fft, noverlap=fft - hop)[1])
    T0, V0 = rng.random((B, C, F, K)), rng.random((B, C, K, T))
    want = BatchedGaussILRMA(n_basis=K).separate_waveform_batch(xf, fft, hop, iteration=5, basis=T0, activation=V0, pipeline=1)
    got = BatchedGaussILRMA(n_basis=K).separate_waveform_batch(pcm, fft, hop, iteration=5, basis=T0, activation=V0, pipeline=2)
    assert got.dtype == np.float32 and got.shape == want.shape
    assert np.array_equal(got, want)
    loss = np.zeros(B)
    BatchedGaussILRMA(n_basis=K).separate_waveform_batch(pcm, fft, hop, iteration=5, basis=T0, activation=V0, pipeline=2, loss_out=loss)
    assert np.all(np.isfinite(loss)) and np.all(loss != 0)
