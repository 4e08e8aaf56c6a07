#!/usr/bin/env python
"""A/B check of the power hand-off between the basis and the activation update (n_basis == 2): the same job in two
processes, one with BSSGPU_NO_POWER_HANDOFF=1, must give bit-identical state (the hand-off moves values, it does not
change arithmetic).  Prints one JSON line."""
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def child(path):
    from audio_source_separation_b200.bss.ilrma import GaussILRMA, tILRMA
    from oracle import synth
    out = {}
    for tag, (C, F, T) in {'c4': (4, 257, 512), 'c2': (2, 129, 200), 'c3': (3, 65, 132)}.items():
        X = synth.mix2(C, F, T, seed=3)
        W0, T0, V0 = synth.initial_state(C, F, T, 2, seed=7)
        m = GaussILRMA(n_basis=2, recordable_loss=False)
        out[tag + '_y'] = m(X, iteration=12, demix_filter=W0, basis=T0, activation=V0)
        out[tag + '_t'] = m.basis
        out[tag + '_v'] = m.activation
    X = synth.mix2(3, 33, 64, seed=5)
    W0, T0, V0 = synth.initial_state(3, 33, 64, 2, seed=7)
    m = tILRMA(n_basis=2, nu=5.0, recordable_loss=False)
    out['t_y'] = m(X, iteration=5, demix_filter=W0, basis=T0, activation=V0)
    out['t_v'] = m.activation
    np.savez(path, **out)


if __name__ == '__main__':
    if len(sys.argv) > 1:
        child(sys.argv[1])
        sys.exit(0)
    with tempfile.TemporaryDirectory() as d:
        files = []
        for off in (False, True):
            env = dict(os.environ)
            env.pop('BSSGPU_NO_POWER_HANDOFF', None)
            if off:
                env['BSSGPU_NO_POWER_HANDOFF'] = '1'
            f = os.path.join(d, 'off.npz' if off else 'on.npz')
            subprocess.check_call([sys.executable, os.path.abspath(__file__), f], env=env)
            files.append(np.load(f))
        a, b = files
        res = {k: bool(np.array_equal(a[k], b[k])) for k in a.files}
        print(json.dumps({'bit_identical': all(res.values()), 'per_tensor': res}))
        sys.exit(0 if all(res.values()) else 1)
