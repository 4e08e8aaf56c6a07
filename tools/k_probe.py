#!/usr/bin/env python
"""ms per iteration of Gauss-ILRMA-IP at a given n_basis (and batch), for the launch list under ncu.
   python tools/k_probe.py K [B]"""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from audio_source_separation_b200 import _lib
K = int(sys.argv[1]) if len(sys.argv) > 1 else 10
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1
C, F, T = 4, 2049, 512
rng = np.random.default_rng(0)
X = (rng.standard_normal((B, C, F, T), dtype=np.float32) + 1j * rng.standard_normal((B, C, F, T), dtype=np.float32)).astype(np.complex64)
h = _lib.Handle(method=_lib.GAUSS_ILRMA, n_batch=B, n_channels=C, n_sources=C, n_bins=F, n_frames=T, n_basis=K)
h.set_input(X); h.reset_spatial()
h.set_state(_lib.STATE_BASIS, rng.random((B, C, F, K)), np.float64)
h.set_state(_lib.STATE_ACTIVATION, rng.random((B, C, K, T)), np.float64)
h.run(14); h.synchronize()
best = 1e9
for _ in range(3):
    h.timer_begin(); h.run(40); best = min(best, h.timer_end() / 40)
print(json.dumps({"K": K, "B": B, "ms_per_iter": round(best, 4), "loss_finite": bool(np.all(np.isfinite(h.loss())))}))
