#!/usr/bin/env python
"""Timeline of the pipelined end-to-end job (BatchedGaussILRMA.separate_batch): per sub-batch, when the upload, the
update loop and the download finish (host clock, seconds since the start of the call)."""
import sys, os, time, threading
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from audio_source_separation_b200 import _lib
from audio_source_separation_b200.batch import shard_range, ramp_sizes
from audio_source_separation_b200._model import parse_spatial, parse_normalize

B, C, F, T, K = 64, 4, 2049, 512, 2
arg = sys.argv[1] if len(sys.argv) > 1 else 'ramp'   # 'ramp', a count, or comma-separated sub-batch sizes
steps = 100
x = torch.empty((B, C, F, T), dtype=torch.complex64, pin_memory=True)
x.numpy()[:] = (np.random.default_rng(0).standard_normal((B, C, F, T), dtype=np.float32)
                + 1j * np.random.default_rng(1).standard_normal((B, C, F, T), dtype=np.float32))
y = torch.empty((B, C, F, T), dtype=torch.complex64, pin_memory=True)
X, Y = x.numpy(), y.numpy()
rng = np.random.default_rng(7)
T0 = rng.random((B, C, F, K)); V0 = rng.random((B, C, K, T))
if arg == 'ramp' or ',' in arg:
    sizes = ramp_sizes(B) if arg == 'ramp' else [int(v) for v in arg.split(',')]
    edges = np.concatenate(([0], np.cumsum(sizes)))
    spans = [(int(edges[i]), int(edges[i + 1])) for i in range(len(sizes))]
else:
    spans = [shard_range(B, i, int(arg)) for i in range(int(arg))]
P = len(spans)
hs = [_lib.Handle(method=_lib.GAUSS_ILRMA, spatial=0, normalize=1, n_batch=hi - lo, n_channels=C, n_sources=C, n_bins=F, n_frames=T,
                  n_basis=K, stream_priority=-(P - 1 - i)) for i, (lo, hi) in enumerate(spans)]
for rep in range(3):
    marks = [dict() for _ in range(P)]
    t0 = time.perf_counter()
    def job(i):
        lo, hi = spans[i]; h = hs[i]; m = marks[i]
        h.reset_spatial()   # same order as batch.py: small uploads first
        h.set_state(_lib.STATE_BASIS, T0[lo:hi], np.float64)
        h.set_state(_lib.STATE_ACTIVATION, V0[lo:hi], np.float64); m['state'] = time.perf_counter() - t0
        h.set_input_ptr(X[lo:hi].ctypes.data, _lib.C64); m['input'] = time.perf_counter() - t0
        h.run(steps); m['queued'] = time.perf_counter() - t0
        h.synchronize(); m['loop'] = time.perf_counter() - t0
        h.separate_into(Y[lo:hi].ctypes.data, _lib.C64, True); m['out'] = time.perf_counter() - t0
    th = [threading.Thread(target=job, args=(i,)) for i in range(P)]
    [t.start() for t in th]; [t.join() for t in th]
    total = time.perf_counter() - t0
    print("rep", rep, "total %.1f ms" % (1e3 * total))
    for i, m in enumerate(marks):
        print("  part", i, {k: round(1e3 * v, 1) for k, v in m.items()})
