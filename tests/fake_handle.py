"""Oracle-backed stand-in for `_lib.Handle` (method BSS_GAUSS_IDLMA) -- TEST INFRASTRUCTURE.

Lets the host-side logic of a drop-in class (attribute protocol, call order, exceptions, what is uploaded when) run in
the CPU suite, where there is no GPU: every device operation is answered by the NumPy oracle in float64.  It is installed
by monkeypatching `audio_source_separation_b200._lib.Handle` inside a test; the product never sees it.
"""
import numpy as np

from audio_source_separation_b200 import _lib
from oracle import core


class FakeIDLMAHandle:
    instances = []

    def __init__(self, **cfg):
        assert cfg['method'] == _lib.GAUSS_IDLMA
        self.cfg = cfg
        self.calls = []
        self.X = self.W = self.R = None
        FakeIDLMAHandle.instances.append(self)

    # data movement
    def set_input(self, x):
        self.calls.append('set_input')
        self.X = np.array(x, dtype=np.complex128)

    def reset_spatial(self):
        self.calls.append('reset_spatial')
        C, F, _ = self.X.shape
        self.W = np.tile(np.eye(C, dtype=np.complex128), (F, 1, 1))

    def set_state(self, which, a, dtype):
        self.calls.append(('set_state', which))
        if which == _lib.STATE_VARIANCE:
            R = np.array(a, dtype=np.float64)
            R[R < self.cfg['eps']] = self.cfg['eps']
            self.R = R
        elif which == _lib.STATE_DEMIX_FILTER:
            self.W = np.array(a, dtype=np.complex128)
        else:
            raise ValueError("state cannot be set")

    def get_state(self, which, shape, dtype):
        if which == _lib.STATE_DEMIX_FILTER:
            return self.W.copy()
        if which == _lib.STATE_ESTIMATION:
            return core.demix(self.X, self.W)
        raise ValueError("unknown state")

    # update loop
    def update_once(self):
        self.calls.append('update_once')
        assert self.R is not None, "GaussIDLMA: set the source variances (dnn_output) first"
        U = core.weighted_covariance(self.X, self.R)
        core.ip_rows(self.W, U, self.cfg['threshold'])

    def _scale(self):
        Y = core.demix(self.X, self.W)
        return Y, core.projection_back_scale(Y, self.X[self.cfg['reference_id']])

    def normalize(self):
        self.calls.append('normalize')
        assert self.cfg['normalize'] == _lib.NORMALIZE_PROJECTION_BACK
        _, scale = self._scale()
        self.W = self.W * scale.T[:, :, np.newaxis]

    def loss(self):
        P = np.abs(core.demix(self.X, self.W)) ** 2
        value = np.sum(P / self.R + np.log(self.R)) - 2 * self.X.shape[2] * core.logabsdet_sum(self.W)
        return np.array([value])

    def separate(self, shape, dtype=np.complex128, projection_back=True):
        Y, scale = self._scale()
        return Y * scale[..., np.newaxis] if projection_back else Y

    def launch_count(self):
        return len(self.calls)

    def close(self):
        self.calls.append('close')


class FakeILRMAHandle:
    """Gauss-ILRMA (non-partitioned; IP, IP2, ISS; 'power' / 'projection-back' / no normalisation) answered by
    oracle/ilrma.py in float64.  Mirrors the call semantics of the real handle: `run` advances the IP2 pair schedule like
    the reference's __call__, ISS carries estimates instead of a filter, `compute_demix_filter` rebuilds W from them."""
    instances = []
    SPATIAL = {_lib.SPATIAL_IP: 'IP', _lib.SPATIAL_ISS: 'ISS', _lib.SPATIAL_IP2: 'IP2'}
    NORMALIZE = {_lib.NORMALIZE_NONE: False, _lib.NORMALIZE_POWER: 'power', _lib.NORMALIZE_PROJECTION_BACK: 'projection-back'}

    def __init__(self, **cfg):
        assert cfg['method'] == _lib.GAUSS_ILRMA and not cfg.get('partitioning')
        self.cfg = cfg
        self.spatial = self.SPATIAL[cfg['spatial']]
        self.norm = self.NORMALIZE[cfg['normalize']]
        self.calls = []
        self.X = self.W = self.Y = self.T = self.V = None
        self.pair = None
        FakeILRMAHandle.instances.append(self)

    def _state(self):
        return {'X': self.X, 'W': None if self.spatial == 'ISS' else self.W, 'Y': self.Y, 'T': self.T, 'V': self.V,
                'pair': self.pair}

    def _take(self, st):
        if st['W'] is not None:
            self.W = st['W']
        self.Y, self.T, self.V = st['Y'], st['T'], st['V']

    def _refresh(self):
        if self.X is not None and self.W is not None:
            self.Y = core.demix(self.X, self.W)

    def set_input(self, x):
        self.calls.append('set_input')
        self.X = np.array(x, dtype=np.complex128)
        self._refresh()

    def reset_spatial(self):
        self.calls.append('reset_spatial')
        C, F, _ = self.X.shape
        self.W = np.tile(np.eye(C, dtype=np.complex128), (F, 1, 1))
        self.pair = None
        self._refresh()

    def set_state(self, which, a, dtype):
        self.calls.append(('set_state', which))
        a = np.array(a, dtype=dtype)
        if which == _lib.STATE_DEMIX_FILTER:
            self.W = a
            self._refresh()
        elif which == _lib.STATE_BASIS:
            self.T = a
        elif which == _lib.STATE_ACTIVATION:
            self.V = a
        else:
            raise ValueError("state cannot be set")

    def get_state(self, which, shape, dtype):
        self.calls.append(('get_state', which))
        value = {_lib.STATE_DEMIX_FILTER: self.W, _lib.STATE_BASIS: self.T, _lib.STATE_ACTIVATION: self.V,
                 _lib.STATE_ESTIMATION: self.Y if self.spatial == 'ISS' else core.demix(self.X, self.W)}[which]
        assert tuple(value.shape) == tuple(shape)
        return np.array(value, dtype=dtype)

    def set_update_pair(self, m, n):
        self.calls.append(('set_update_pair', m, n))
        self.pair = None if (m, n) == (-1, -1) else (int(m), int(n))

    def update_once(self):
        from oracle import ilrma as o_ilrma
        self.calls.append('update_once')
        st = self._state()
        o_ilrma.update_once(st, self.spatial, self.cfg['domain'], self.norm, False, self.cfg['reference_id'], self.cfg['eps'],
                            self.cfg['threshold'])
        self._take(st)

    def _advance_pair(self):
        if self.spatial == 'IP2':
            self.pair = core.next_update_pair(self.pair, self.X.shape[0])

    def run(self, n_iter):
        self.calls.append(('run', n_iter))
        for _ in range(n_iter):
            self._advance_pair()
            self.update_once()

    def run_record(self, n_iter):
        self.calls.append(('run_record', n_iter))
        out = np.empty((n_iter, 1))
        for i in range(n_iter):
            self._advance_pair()
            self.update_once()
            out[i, 0] = self.loss()[0]
        return out

    def loss(self):
        from oracle import ilrma as o_ilrma
        return np.array([o_ilrma.negative_loglikelihood(self._state(), self.cfg['domain'], False, self.cfg['eps'])])

    def compute_demix_filter(self):
        self.calls.append('compute_demix_filter')
        self.W = core.estimate_demix_filter(self.Y, self.X)

    def separate(self, shape, dtype=np.complex128, projection_back=True):
        Y = self.Y if self.spatial == 'ISS' else core.demix(self.X, self.W)
        if projection_back:
            Y = Y * core.projection_back_scale(Y, self.X[self.cfg['reference_id']])[..., np.newaxis]
        return np.array(Y, dtype=dtype)

    def launch_count(self):
        return len(self.calls)

    def close(self):
        self.calls.append('close')


class FakeAuxIVAHandle(FakeILRMAHandle):
    """AuxLaplaceIVA / AuxGaussIVA (IP, IP2, ISS) answered by oracle/auxiva.py; same call semantics as above, no source model."""
    instances = []

    def __init__(self, **cfg):
        assert cfg['method'] in (_lib.AUX_LAPLACE_IVA, _lib.AUX_GAUSS_IVA)
        self.cfg = cfg
        self.kind = 'laplace' if cfg['method'] == _lib.AUX_LAPLACE_IVA else 'gauss'
        self.spatial = self.SPATIAL[cfg['spatial']]
        self.calls = []
        self.X = self.W = self.Y = self.T = self.V = None
        self.pair = None
        FakeAuxIVAHandle.instances.append(self)

    def update_once(self):
        from oracle import auxiva as o_auxiva
        self.calls.append('update_once')
        st = self._state()
        o_auxiva.update_once(st, self.kind, self.spatial, self.cfg['eps'], self.cfg['threshold'])
        self._take(st)

    def loss(self):
        from oracle import auxiva as o_auxiva
        return np.array([o_auxiva.negative_loglikelihood(self._state(), self.kind, self.cfg['eps'])])
