"""Where does the time of one drop-in call go?  (single cfg3 mixture, default recordable_loss=True)"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from audio_source_separation_b200.bss.ilrma import GaussILRMA
from audio_source_separation_b200 import _lib
from oracle import synth
X = synth.mix2(4, 2049, 512, seed=0)
W0, T0, V0 = synth.initial_state(4, 2049, 512, 2, seed=7)
for rep in range(3):
    t0 = time.perf_counter()
    m = GaussILRMA(n_basis=2)
    m.input = X
    m._reset(demix_filter=W0, basis=T0, activation=V0)
    t1 = time.perf_counter()
    l0 = m.compute_negative_loglikelihood()
    t2 = time.perf_counter()
    loss = m._handle.run_record(100)
    t3 = time.perf_counter()
    out = m._handle.separate((4, 2049, 512), np.complex128, projection_back=True)
    t4 = time.perf_counter()
    print("rep %d: reset %.1f ms, first loss %.1f ms, 100 x (update + loss) %.1f ms, separate + D2H %.1f ms" % (
        rep, 1e3 * (t1 - t0), 1e3 * (t2 - t1), 1e3 * (t3 - t2), 1e3 * (t4 - t3)))
    t0 = time.perf_counter()
    out = GaussILRMA(n_basis=2)(X, iteration=100, demix_filter=W0, basis=T0, activation=V0)
    print("   whole __call__: %.1f ms" % (1e3 * (time.perf_counter() - t0)))
