#!/usr/bin/env python
"""One pipelined waveform job (BatchedGaussILRMA.separate_waveform_batch) on one GPU: host-side phase marks of every sub-batch,
for `pipeline` given on the command line.  Run under `ncu --metrics gpu__time_duration.sum` for the per-kernel device times."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from audio_source_separation_b200.batch import BatchedGaussILRMA, ramp_sizes

B = int(os.environ.get('PROBE_B', '64'))
steps = int(os.environ.get('PROBE_STEPS', '20'))
arg = sys.argv[1] if len(sys.argv) > 1 else 'ramp'
pipeline = ramp_sizes(B) if arg == 'ramp' else ([int(v) for v in arg.split(',')] if ',' in arg else int(arg))
C, F, T, K, FFT, HOP = 4, 2049, 512, 2, 4096, 2048
n = (T - 1) * HOP
x = torch.empty((B, C, n), dtype=torch.float32, pin_memory=True)
rng = np.random.default_rng(0)
for b in range(B):
    x.numpy()[b] = rng.standard_normal((C, n), dtype=np.float32)
y = torch.empty((B, C, n), dtype=torch.float32, pin_memory=True)
T0, V0 = rng.random((B, C, F, K)), rng.random((B, C, K, T))
m = BatchedGaussILRMA(n_basis=K)
for rep in range(3):
    t0 = time.perf_counter()
    m.separate_waveform_batch(x.numpy(), FFT, HOP, out=y.numpy(), iteration=steps, basis=T0, activation=V0, pipeline=pipeline)
    dt = time.perf_counter() - t0
    print(json.dumps({"rep": rep, "ms": round(1e3 * dt, 2), "pipeline": pipeline, "timeline": m.timeline}))
