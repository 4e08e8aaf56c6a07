#!/usr/bin/env python
"""Where does the GPU trajectory leave the oracle's on the sample recording?  rel errors of T, V, W, TV per iteration."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
from conftest import GOLDEN, rel
from oracle import ilrma as o_ilrma
from scipy import signal as ss
from audio_source_separation_b200 import _lib

z = np.load(os.path.join(GOLDEN, 'audio_sample2_pcm.npz'))
x = z['pcm'].astype(np.float64) / 32768
_, _, X = ss.stft(x, nperseg=4096, noverlap=2048)
X = X.astype(np.complex64).astype(np.complex128)
C, F, T = X.shape
for K in (int(a) for a in (sys.argv[1:] or ['5', '2'])):
    np.random.seed(111)
    st = o_ilrma.init_state(X, K)
    T0, V0 = st['T'].copy(), st['V'].copy()
    h = _lib.Handle(method=_lib.GAUSS_ILRMA, spatial=_lib.SPATIAL_IP, normalize=_lib.NORMALIZE_POWER, n_batch=1, n_channels=C, n_sources=C,
                    n_bins=F, n_frames=T, n_basis=K)
    h.set_input(X[None]); h.reset_spatial()
    h.set_state(_lib.STATE_BASIS, T0[None], np.float64); h.set_state(_lib.STATE_ACTIVATION, V0[None], np.float64)
    print("K =", K)
    for it in range(1, 101):
        h.update_once(); o_ilrma.update_once(st)
        if it <= 6 or it in (10, 20, 50, 100):
            Tg = h.get_state(_lib.STATE_BASIS, (1, C, F, K), np.float64)[0]
            Vg = h.get_state(_lib.STATE_ACTIVATION, (1, C, K, T), np.float64)[0]
            Wg = h.get_state(_lib.STATE_DEMIX_FILTER, (1, F, C, C), np.complex128)[0]
            r = Tg / st['T']
            print("  it %3d relT %.2e relV %.2e relW %.2e relTV %.2e  T ratio pct(1,50,99) %s  per-source median %s" % (
                it, rel(Tg, st['T']), rel(Vg, st['V']), rel(Wg, st['W']), rel(Tg @ Vg, st['T'] @ st['V']),
                np.round(np.percentile(r, [1, 50, 99]), 5), np.round(np.median(r, axis=(1, 2)), 6)))
    h.close()
