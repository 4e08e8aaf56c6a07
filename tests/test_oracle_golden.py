"""CPU: the oracle restatement reproduces every fixture generated from the unmodified reference."""
import numpy as np
import pytest

from conftest import load_golden, rel
from oracle import core, ilrma, auxiva, fastmnmf, nmf

TOL = 1e-9

ILRMA_CASES = ['ilrma_ip_power_d2', 'ilrma_ip_power_d1', 'ilrma_ip_pb_d2', 'ilrma_iss_power_d2', 'ilrma_iss_pb_d1',
               'ilrma_ip2_power_d2', 'ilrma_ip2_power_c2', 'ilrma_ip_power_part']
AUXIVA_CASES = ['auxiva_laplace_ip', 'auxiva_laplace_ip_c4', 'auxiva_gauss_ip', 'auxiva_laplace_iss', 'auxiva_gauss_iss',
                'auxiva_laplace_ip2']
NMF_CASES = ['nmf_euc_d2', 'nmf_euc_d1', 'nmf_kl_d2', 'nmf_kl_d1', 'nmf_is_mm_d2', 'nmf_is_mm_d1', 'nmf_is_me', 'nmf_t',
             'nmf_cauchy_naive_multipricative', 'nmf_cauchy_mm', 'nmf_cauchy_me', 'nmf_cauchy_mm_fast']


@pytest.mark.parametrize('name', ILRMA_CASES)
def test_ilrma(name):
    meta, i, o = load_golden(name)
    presets = {'W': i['W0'], 'T': i['T0'], 'V': i['V0']}
    if meta['partitioning']:
        presets['Z'] = i['Z0']
    out, st, loss = ilrma.run(i['X'], iteration=meta['iteration'], n_basis=meta['n_basis'], spatial=meta['algorithm_spatial'],
                              domain=meta['domain'], normalize_mode=meta['normalize'], partitioning=meta['partitioning'], **presets)
    tol = 1e-7 if meta['algorithm_spatial'] in ('IP2', 'pairwise') else TOL
    assert rel(out, o['output']) < tol
    assert rel(st['T'], o['basis']) < tol
    assert rel(st['V'], o['activation']) < tol
    assert rel(loss, o['loss']) < tol
    W = st['W'] if st['W'] is not None else st['W_final']
    assert rel(W, o['demix_filter']) < tol


def test_ilrma_seeded_dropin():
    meta, i, o = load_golden('ilrma_seeded_dropin')
    np.random.seed(meta['seed'])
    out, st, loss = ilrma.run(i['X'], iteration=meta['iteration'], n_basis=meta['n_basis'])
    assert rel(out, o['output']) < TOL and rel(loss, o['loss']) < TOL


def test_tilrma():
    meta, i, o = load_golden('tilrma_nu5')
    out, st, loss = ilrma.t_run(i['X'], iteration=meta['iteration'], n_basis=meta['n_basis'], nu=meta['nu'], W=i['W0'], T=i['T0'],
                                V=i['V0'])
    assert rel(out, o['output']) < TOL and rel(loss, o['loss']) < TOL and rel(st['W'], o['demix_filter']) < TOL


@pytest.mark.parametrize('name', AUXIVA_CASES)
def test_auxiva(name):
    meta, i, o = load_golden(name)
    kind = 'laplace' if meta['model'] == 'AuxLaplaceIVA' else 'gauss'
    out, st, loss = auxiva.run(i['X'], iteration=meta['iteration'], kind=kind, spatial=meta['algorithm_spatial'], W=i['W0'])
    tol = 1e-7 if meta['algorithm_spatial'] in ('IP2', 'pairwise') else TOL
    assert rel(out, o['output']) < tol and rel(loss, o['loss']) < tol


@pytest.mark.parametrize('name', ['fastmnmf_m3n3', 'fastmnmf_m4n2'])
def test_fastmnmf(name):
    meta, i, o = load_golden(name)
    out, st, loss = fastmnmf.run(i['X'], iteration=meta['iteration'], n_basis=meta['n_basis'], n_sources=meta['n_sources'],
                                 W=i['W0'], H=i['H0'])
    assert rel(out, o['output']) < TOL and rel(loss, o['loss']) < TOL
    assert rel(st['Q'], o['diagonalizer']) < TOL and rel(st['G'], o['spatial_covariance']) < TOL


@pytest.mark.parametrize('name', NMF_CASES)
def test_nmf(name):
    meta, i, o = load_golden(name)
    T, V, loss = nmf.run(meta['model'], i['Z'], n_basis=meta['n_basis'], iteration=meta['iteration'], domain=meta['domain'],
                         algorithm=meta['algorithm'], nu=meta['nu'], T=i['T0'], V=i['V0'])
    assert rel(T, o['basis']) < TOL and rel(V, o['activation']) < TOL and rel(loss, o['loss']) < TOL


def test_primitives():
    meta, i, o = load_golden('primitives')
    assert rel(core.weighted_covariance(i['X'], i['R']), o['U']) < 1e-13
    Y = core.demix(i['X'], i['W'])
    assert rel(core.projection_back_scale(Y, i['X'][0]), o['scale2']) < 1e-13
    assert rel(core.projection_back_scale(Y, i['X']), o['scale3']) < 1e-13
    order = np.argsort(i['eigval'], axis=-1)[:, ::-1]
    assert np.array_equal(order, o['order'])
    assert np.array_equal(core.gather_by_order(i['eigvec'].swapaxes(-2, -1), order, axis=-2), o['sorted'])
    for n_src in (2, 3, 4):
        pair, seq = None, []
        for _ in range(7):
            pair = core.next_update_pair(pair, n_src)
            seq.append(pair)
        assert np.array_equal(np.array(seq), o['pairs{}'.format(n_src)])


def test_known_answer_losses():
    """first loss values of the reference on mix2(4,513,128), K=2 (unrounded initial state)."""
    from oracle import synth
    X = synth.mix2(4, 513, 128, seed=0)
    W0, T0, V0 = synth.initial_state(4, 513, 128, 2, seed=7, round32=False)
    _, _, loss = ilrma.run(X, iteration=2, n_basis=2, W=W0, T=T0, V=V0)
    want = [7.2776542359e5, 4.3845892818e4, 2.8737622949e4]
    assert np.allclose(loss, want, rtol=1e-9)


def test_consistent_ilrma_is_projection_back_ilrma():
    """ConsistentGaussILRMA (src/bss/ilrma.py:1102-1233): the STFT-consistency projection never reaches the IP update, so
    the reference's result equals Gauss-ILRMA with projection-back normalisation (fixture generated from the reference)."""
    meta, i, o = load_golden('ilrma_consistent')
    out, st, loss = ilrma.run(i['X'], iteration=meta['iteration'], n_basis=meta['n_basis'], spatial='IP', domain=2,
                              normalize_mode='projection-back', W=i['W0'], T=i['T0'], V=i['V0'])
    assert rel(out, o['output']) < TOL and rel(loss, o['loss']) < TOL and rel(st['W'], o['demix_filter']) < TOL


def test_stft_fixture_matches_scipy():
    """The reference's stft/istft are thin wrappers of scipy.signal (src/transform/stft.py:4-17): the fixture written from
    the reference must equal a direct scipy call, which is what the GPU tests use as the run-time oracle."""
    from scipy import signal as ss
    meta, i, o = load_golden('ilrma_consistent')
    n, h = meta['fft_size'], meta['hop_size']
    Z = ss.stft(i['x'], nperseg=n, noverlap=n - h, window='hann')[2]
    assert rel(Z, o['stft']) < 1e-14
    y = ss.istft(Z, nperseg=n, noverlap=n - h, window='hann')[1][..., :meta['n_samples']]
    assert rel(y, o['istft']) < 1e-12


MNMF_SAWADA_CASES = ['mnmf_sawada_c2n2', 'mnmf_sawada_c3n2', 'mnmf_sawada_c4n3_eye']


@pytest.mark.parametrize('name', MNMF_SAWADA_CASES)
@pytest.mark.parametrize('closed_form', [False, True])
def test_mnmf_sawada(name, closed_form):
    """Sawada IS-MNMF against the reference fixtures, with the reference's 2M x 2M eigen Riccati solver and with the
    Hermitian closed form the CUDA path uses."""
    from oracle import mnmf
    meta, i, o = load_golden(name)
    riccati = mnmf.solve_riccati_hermitian if closed_form else mnmf.solve_riccati
    out, st, loss = mnmf.run(i['X'], iteration=meta['iteration'], n_basis=meta['n_basis'], n_sources=meta['n_sources'],
                             normalize=meta['normalize'], riccati=riccati, H=i['H0'], Z=i['Z0'], T=i['T0'], V=i['V0'])
    tol = 1e-8
    assert rel(out, o['output']) < tol
    assert rel(st['H'], o['spatial']) < tol
    assert rel(st['Z'], o['latent']) < tol and rel(st['T'], o['basis']) < tol and rel(st['V'], o['activation']) < tol
    assert rel(loss, o['loss']) < tol


@pytest.mark.parametrize('name', ['idlma_gauss_d2', 'idlma_gauss_d1'])
def test_idlma(name):
    """GaussIDLMA (src/sss/idlma.py) with the toy DNN; the DNN runs in torch float32 on this machine's CPU, so 1e-6."""
    from oracle import idlma, synth
    meta, i, o = load_golden(name)
    dnn = synth.dnn_as_callable(synth.toy_dnn())
    out, st, loss = idlma.run(i['X'], dnn, iteration=meta['iteration'], domain=meta['domain'])
    assert rel(out, o['output']) < 1e-6 and rel(st['W'], o['demix_filter']) < 1e-6
    assert rel(st['dnn_output'], o['dnn_output']) < 1e-6 and rel(loss, o['loss']) < 1e-6
    st2 = idlma.init_state(i['X'])
    st2['dnn_output'] = i['R0'].copy()
    idlma.update_space_model(st2, meta['domain'])
    assert rel(st2['W'], o['space_W1']) < TOL
    assert abs(idlma.negative_loglikelihood(st2, meta['domain']) - o['space_loss1']) < 1e-6 * abs(o['space_loss1'])


def test_every_fixture_is_checked_on_both_sides():
    """Each committed fixture (generated from the unmodified reference) is consumed by a CPU test of the oracle and by a GPU
    parity test -- a fixture nobody reads would be a silent gap in the parity claim."""
    import glob
    import os
    here = os.path.dirname(os.path.abspath(__file__))
    names = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(here, 'golden', '*.npz')))
    assert len(names) >= 37
    cpu_src = open(os.path.join(here, 'test_oracle_golden.py')).read()
    gpu_src = ''.join(open(p).read() for p in glob.glob(os.path.join(here, 'test_gpu_*.py')))

    def used(name, src):
        # fixtures are referenced by full name or built from a family prefix ('nmf_cauchy_' + alg, 'nmf_' + kind ...)
        return name in src or any(name.startswith(pre) and ("'" + pre + "'") in src for pre in ('nmf_cauchy_', 'nmf_'))

    assert [n for n in names if not used(n, cpu_src)] == []
    assert [n for n in names if not used(n, gpu_src)] == []


@pytest.mark.parametrize('name', ['audio_sample2_auxiva_laplace_ip', 'audio_sample2_ilrma_k5'])
def test_oracle_reproduces_the_reference_on_the_sample_recording(name):
    """100 iterations on the reference's own sample recording (cond_2 up to 3.4e11): the fixtures hold what the unmodified
    reference produced (oracle/pin/make_golden.py audio); the oracle must land on the same trajectory."""
    import os
    from scipy import signal as ss
    from conftest import GOLDEN
    meta, _, o = load_golden(name)
    assert meta['pcm_file'] == 'audio_sample2_pcm'    # the int16 samples of dataset/sample-song/sample-2_mixture_16000.wav
    z = np.load(os.path.join(GOLDEN, meta['pcm_file'] + '.npz'))
    x = z['pcm'].astype(np.float64) / 32768
    _, _, X = ss.stft(x, nperseg=meta['fft_size'], noverlap=meta['fft_size'] - meta['hop_size'])
    assert X.shape == (2, 2049, 209)
    half = int(o['half'])
    if meta['model'] == 'AuxLaplaceIVA':
        from oracle import auxiva
        _, st_half, _ = auxiva.run(X, iteration=half, kind='laplace', record_loss=False)
        out, st, loss = auxiva.run(X, iteration=meta['iteration'], kind='laplace')
    else:
        from oracle import ilrma
        np.random.seed(meta['seed'])
        _, st_half, _ = ilrma.run(X, iteration=half, n_basis=meta['n_basis'], record_loss=False)
        assert rel(st_half['T'], o['basis_half']) < 1e-7 and rel(st_half['V'], o['activation_half']) < 1e-7
        np.random.seed(meta['seed'])
        out, st, loss = ilrma.run(X, iteration=meta['iteration'], n_basis=meta['n_basis'])
        assert rel(st['T'], o['basis']) < 1e-6 and rel(st['V'], o['activation']) < 1e-6
    assert rel(st_half['W'], o['demix_filter_half']) < 1e-7
    assert rel(out[:, ::meta['bin_step']], o['output_bins']) < 1e-6       # the fixture stores these bins as complex64
    assert rel(np.linalg.norm(out, axis=2), o['output_bin_norms']) < 1e-6
    assert rel(st['W'], o['demix_filter']) < 1e-6 and rel(loss, o['loss']) < 1e-9
