import sys, os
sys.path.insert(0, '/root/repo')
import numpy as np
from audio_source_separation_b200.bss.ilrma import GaussILRMA
from oracle import synth
X = synth.mix2(4, 2049, 512, seed=0)
W0, T0, V0 = synth.initial_state(4, 2049, 512, 2, seed=7)
os.environ['BSSGPU_NO_GRAPH'] = '1'
m = GaussILRMA(n_basis=2, recordable_loss=False)
m.input = X
m._reset(demix_filter=W0, basis=T0, activation=V0)
m._handle.run(6)
m._handle.synchronize()
