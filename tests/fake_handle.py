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
