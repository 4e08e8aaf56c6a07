#!/usr/bin/env python
"""A/B of the source-model forms on the 64-mixture shard: ms per iteration of the device loop.
   BSSGPU_FUSED_CTAS=2|3|4 python tools/ab_fused.py fused     /    python tools/ab_fused.py three"""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from audio_source_separation_b200 import _lib
mode = sys.argv[1] if len(sys.argv) > 1 else 'fused'
B, C, F, T, K = int(os.environ.get('AB_B', '64')), 4, 2049, 512, 2
rng = np.random.default_rng(0)
X = np.empty((B, C, F, T), dtype=np.complex64)
for b in range(B):
    X[b] = (rng.standard_normal((C, F, T), dtype=np.float32) + 1j * rng.standard_normal((C, F, T), dtype=np.float32)) * (0.2 + rng.random((C, F, 1), dtype=np.float32))
h = _lib.Handle(method=_lib.GAUSS_ILRMA, spatial=_lib.SPATIAL_IP, normalize=_lib.NORMALIZE_POWER, n_batch=B, n_channels=C, n_sources=C, n_bins=F,
                n_frames=T, n_basis=K)
h.set_option(_lib.OPT_SOURCE_MODEL, _lib.SOURCE_MODEL_FUSED if mode == 'fused' else _lib.SOURCE_MODEL_THREE_PASS)
h.set_input(X); h.reset_spatial()
h.set_state(_lib.STATE_BASIS, rng.random((B, C, F, K)), np.float64)
h.set_state(_lib.STATE_ACTIVATION, rng.random((B, C, K, T)), np.float64)
h.run(14); h.synchronize()
best = 1e9
for _ in range(3):
    h.timer_begin(); h.run(40); best = min(best, h.timer_end() / 40)
print(json.dumps({"mode": mode, "fused_ctas": os.environ.get('BSSGPU_FUSED_CTAS'), "B": B, "ms_per_iter": round(best, 4),
                  "source_model": h.get_info(_lib.INFO_SOURCE_MODEL), "chunks": h.get_info(_lib.INFO_ACT_CHUNKS), "loss_finite": bool(np.all(np.isfinite(h.loss())))}))
