import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from audio_source_separation_b200 import _lib
B, C, F, T, K = 16, 4, 2049, 512, 10
rng = np.random.default_rng(0)
X = (rng.standard_normal((B, C, F, T), dtype=np.float32) + 1j * rng.standard_normal((B, C, F, T), dtype=np.float32)).astype(np.complex64)
os.environ['BSSGPU_NO_GRAPH'] = '1'
h = _lib.Handle(method=_lib.GAUSS_ILRMA, n_batch=B, n_channels=C, n_sources=C, n_bins=F, n_frames=T, n_basis=K)
h.set_input(X); h.reset_spatial()
h.set_state(_lib.STATE_BASIS, rng.random((B, C, F, K)), np.float64)
h.set_state(_lib.STATE_ACTIVATION, rng.random((B, C, K, T)), np.float64)
h.run(4); h.synchronize()
