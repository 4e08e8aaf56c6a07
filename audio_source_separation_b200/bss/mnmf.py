"""FastMNMF on the GPU with the reference's class surface (src/bss/mnmf.py).

`FastMultichannelISNMF` (src/bss/mnmf.py:637-946): same constructor arguments, `__call__`, `update_once`
(`update_NMF`, `update_SCM`, `update_diagonalizer` and the three-step power normalisation), `separate`
(multichannel Wiener filter through the inverse diagonaliser, reference-microphone image),
`compute_negative_loglikelihood`, callbacks and the public state attributes `basis`, `activation`,
`diagonalizer`, `spatial_covariance`, `estimation`.  As in the reference, `_reset` re-creates `diagonalizer`
and `spatial_covariance` unconditionally (:660-663, :688-689); only `basis` / `activation` can be preset
through kwargs.  The Sawada MM / Ozerov EM `MultichannelISNMF` (:115-635) is not part of this package.
"""
import numpy as np

from .. import _lib
from .._model import DeviceModel, parse_normalize

EPS = 1e-12
THRESHOLD = 1e+12


class MultichannelNMFbase(DeviceModel):
    """src/bss/mnmf.py:25-113"""

    _STATE_IDS = {'basis': _lib.STATE_BASIS, 'activation': _lib.STATE_ACTIVATION, 'diagonalizer': _lib.STATE_DIAGONALIZER,
                  'spatial_covariance': _lib.STATE_SPATIAL, 'estimation': _lib.STATE_ESTIMATION}

    def __init__(self, n_basis=10, n_sources=None, callbacks=None, recordable_loss=True, eps=EPS):
        """
        Args:
            n_basis: number of basis
        """
        DeviceModel.__init__(self)
        if callbacks is not None:
            if callable(callbacks):
                callbacks = [callbacks]
            self.callbacks = callbacks
        else:
            self.callbacks = None

        self.eps = eps
        self.n_basis = n_basis
        self.n_sources = n_sources

        self.input = None
        self.recordable_loss = recordable_loss
        if self.recordable_loss:
            self.loss = []
        else:
            self.loss = None

    def _reset(self, **kwargs):
        assert self.input is not None, "Specify data!"

        for key in kwargs.keys():
            setattr(self, key, kwargs[key])

        n_sources = self.n_sources

        X = self.input
        n_channels, n_bins, n_frames = X.shape

        if n_sources is None:
            n_sources = n_channels
        self.n_sources, self.n_channels = n_sources, n_channels
        self.n_bins, self.n_frames = n_bins, n_frames

    def update_once(self):
        raise NotImplementedError("Implement 'update_once' method")

    def separate(self, input):
        raise NotImplementedError("Implement 'update_once' method")

    def compute_negative_loglikelihood(self):
        raise NotImplementedError("Implement 'compute_negative_loglikelihood' method.")


class FastMultichannelISNMF(MultichannelNMFbase):
    """
    Reference: "Fast Multichannel Source Separation Based on Jointly Diagonalizable Spatial Covariance Matrices"
    Drop-in for src/bss/mnmf.py:637-946.
    """

    def __init__(self, n_basis=10, n_sources=None, partitioning=False, normalize='power', reference_id=0, callbacks=None,
                 recordable_loss=True, eps=EPS, threshold=THRESHOLD):
        super().__init__(n_basis=n_basis, n_sources=n_sources, callbacks=callbacks, recordable_loss=recordable_loss, eps=eps)

        self.partitioning = partitioning
        self.normalize = normalize
        self.reference_id = reference_id

        self.threshold = threshold

    # -- device plumbing -----------------------------------------------------------------------------
    def _state_shape(self, name):
        N, M, F, T, K = self.n_sources, self.n_channels, self.n_bins, self.n_frames, self.n_basis
        if name == 'basis':
            return (N, F, K)
        if name == 'activation':
            return (N, K, T)
        if name == 'diagonalizer':
            return (F, M, M)
        if name == 'spatial_covariance':
            return (N, F, M)
        if name == 'estimation':
            return (N, F, T)
        raise KeyError(name)

    def _normalize_code(self):
        if not self.normalize:
            return _lib.NORMALIZE_NONE
        if self.normalize != 'power':
            raise ValueError("Not support normalization based on {}. Choose 'power'".format(self.normalize))
        return _lib.NORMALIZE_POWER

    def _config(self):
        return dict(method=_lib.FAST_MNMF, normalize=self._normalize_code(), n_batch=1, n_channels=self.n_channels,
                    n_sources=self.n_sources, n_bins=self.n_bins, n_frames=self.n_frames, n_basis=self.n_basis,
                    reference_id=self.reference_id, eps=float(self.eps), threshold=float(self.threshold))

    def _prepare(self):
        X = self.input
        assert X is not None, "Specify data!"
        if self.partitioning:
            raise ValueError("Not support partitioning function.")
        cfg = self._config()
        self._open_handle(tuple(sorted(cfg.items())), **cfg)
        self._send_input(X)
        self._push()

    # -- reference surface -----------------------------------------------------------------------------
    def _reset(self, **kwargs):
        super()._reset(**kwargs)

        n_bins, n_frames = self.n_bins, self.n_frames
        n_sources = self.n_sources
        n_basis = self.n_basis

        if self.partitioning:
            if not hasattr(self, 'latent'):
                self.latent = np.ones((n_sources, n_basis), dtype=np.float64) / n_sources
            else:
                self.latent = self.latent.copy()
            if not hasattr(self, 'basis'):
                self.basis = np.random.rand(n_bins, n_basis)
            else:
                self.basis = self.basis.copy()
            if not hasattr(self, 'activation'):
                self.activation = np.random.rand(n_basis, n_frames)
            else:
                self.activation = self.activation.copy()
            raise ValueError("Not support partitioning function.")   # every update of the reference raises this (:785, :829)
        if not hasattr(self, 'basis'):
            self.basis = np.random.rand(n_sources, n_bins, n_basis)
        else:
            self.basis = np.array(self.basis, dtype=np.float64, copy=True)
        if not hasattr(self, 'activation'):
            self.activation = np.random.rand(n_sources, n_basis, n_frames)
        else:
            self.activation = np.array(self.activation, dtype=np.float64, copy=True)

        # Q = I and g are re-created on the device whatever the host holds (src/bss/mnmf.py:660-663)
        for name in ('diagonalizer', 'spatial_covariance', 'estimation'):
            self._host.pop(name, None)
            self._dirty.discard(name)
            self._snap.pop(name, None)
        cfg = self._config()
        self._open_handle(tuple(sorted(cfg.items())), **cfg)
        self._send_input(self.input)
        self._dirty.discard('diagonalizer')
        self._dirty.discard('spatial_covariance')
        self._handle.reset_spatial()
        self._push()
        self._on_device.update(('basis', 'activation', 'diagonalizer', 'spatial_covariance', 'estimation'))

    def __call__(self, input, iteration=100, **kwargs):
        """
        Args:
            input (n_channels, n_bins, n_frames)
        Returns:
            output (n_sources, n_bins, n_frames)
        """
        self.input = input

        self._reset(**kwargs)

        if self.recordable_loss:
            loss = self.compute_negative_loglikelihood()
            self.loss.append(loss)

        if self.callbacks is None:
            # nothing observes the intermediate states: the loop (and its loss history) stays on the device
            self._push()
            if self.recordable_loss:
                self.loss.extend(float(v) for v in self._handle.run_record(iteration)[:, 0])
            else:
                self._handle.run(iteration)
            self._device_changed()
        else:
            for idx in range(iteration):
                self.update_once()

                if self.recordable_loss:
                    loss = self.compute_negative_loglikelihood()
                    self.loss.append(loss)

                self.estimation = self.separate(self.input)
                for callback in self.callbacks:
                    callback(self)

        output = self.separate(input)
        self.estimation = output

        return output

    def __repr__(self):
        s = "FastMNMF("
        s += "n_basis={n_basis}"
        if hasattr(self, 'n_sources'):
            s += ", n_sources={n_sources}"
        if hasattr(self, 'n_channels'):
            s += ", n_channels={n_channels}"
        s += ", partitioning={partitioning}"
        s += ", normalize={normalize}"
        s += ")"

        return s.format(**self.__dict__)

    def update_once(self):
        self._prepare()
        self._handle.update_once()
        self._device_changed()

    def compute_negative_loglikelihood(self):
        self._prepare()
        return float(self._handle.loss()[0])

    def separate(self, input):
        """Multichannel Wiener filter with the current model; `input` must be the mixture the model holds
        (src/bss/mnmf.py:919-946 is only ever called that way)."""
        if input is not self.input:
            if self.input is None or np.shape(input) != np.shape(self.input) or not np.array_equal(input, self.input):
                raise NotImplementedError("separate() is available for the model's own input")
        self._prepare()
        return self._handle.separate((self.n_sources, self.n_bins, self.n_frames), np.complex128, projection_back=False)


FastMNMF = FastMultichannelISNMF
