import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from audio_source_separation_b200.algorithm.nmf import EUCNMF
from oracle import synth
Z = synth.spectrogram(257, 128, seed=0)
np.random.seed(111)
m = EUCNMF(n_basis=4)
m.target = Z
m._reset()
m._prepare()
m._handle.run(50)
m._handle.synchronize()
