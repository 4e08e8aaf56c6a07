"""Multichannel NMF on the GPU with the reference's class surface (src/bss/mnmf.py).

`MultichannelISNMF` (src/bss/mnmf.py:116-635, author='Sawada'): the four multiplicative updates of Sawada's MM algorithm
(`update_basis_sawada`, `update_activation_sawada`, `update_latent_sawada`, `update_spatial_sawada` with the Riccati
solve), `reconstruct_covariance`-based log-det loss and the multichannel Wiener filter, in fp64 on the device.  Ozerov's
EM variant is "in progress" upstream and is not provided.

`FastMultichannelISNMF` (src/bss/mnmf.py:637-946): same constructor arguments, `__call__`, `update_once`
(`update_NMF`, `update_SCM`, `update_diagonalizer` and the three-step power normalisation), `separate`
(multichannel Wiener filter through the inverse diagonaliser, reference-microphone image),
`compute_negative_loglikelihood`, callbacks and the public state attributes `basis`, `activation`,
`diagonalizer`, `spatial_covariance`, `estimation`.  As in the reference, `_reset` re-creates `diagonalizer`
and `spatial_covariance` unconditionally (:660-663, :688-689); only `basis` / `activation` can be preset
through kwargs.
"""
import warnings

import numpy as np

from .. import _lib
from .._model import DeviceModel, parse_normalize

EPS = 1e-12
THRESHOLD = 1e+12

__authors__ = ['sawada', 'ozerov']

__kwargs_sawada_mnmf___ = {
    'reference_id': 0
}


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


class MultichannelISNMF(MultichannelNMFbase):
    """
    References:
        Sawada's MNMF: "Multichannel Extensions of Non-Negative Matrix Factorization With Complex-Valued Data"
    Drop-in for src/bss/mnmf.py:116-635 with author='Sawada'.  State attributes: `spatial` (n_bins, n_sources, n_channels,
    n_channels), `latent` (n_sources, n_basis), `basis` (n_bins, n_basis), `activation` (n_basis, n_frames), `estimation`.
    The reference's `covariance_input` (x x^H for every bin and frame, :222-223) is never materialised.
    """

    _STATE_IDS = {'basis': _lib.STATE_BASIS, 'activation': _lib.STATE_ACTIVATION, 'latent': _lib.STATE_LATENT,
                  'spatial': _lib.STATE_SPATIAL, 'estimation': _lib.STATE_ESTIMATION}
    _SNAPSHOT = ('basis', 'activation', 'latent', 'spatial')

    def __init__(self, n_basis=10, n_sources=None, normalize=True, callbacks=None, reference_id=0, author='Sawada',
                 recordable_loss=True, eps=EPS, **kwargs):
        """
        Args:
            n_basis
            n_sources
            normalize
            callbacks <callable> or <list<callable>>: Callback function. Default: None
            reference_id <int>
            author <str>: 'Sawada' ('Ozerov' is in progress upstream and not available here)
            eps <float>: Machine epsilon
        """
        super().__init__(n_basis=n_basis, n_sources=n_sources, callbacks=callbacks, recordable_loss=recordable_loss, eps=eps)

        self.normalize = normalize

        assert author.lower() in __authors__, "Choose from {}".format(__authors__)

        self.author = author

        if author.lower() == 'sawada':
            if set(kwargs) - set(__kwargs_sawada_mnmf___) != set():
                raise ValueError("Invalid keywords.")
            for key in __kwargs_sawada_mnmf___.keys():
                setattr(self, key, __kwargs_sawada_mnmf___[key])
            for key in kwargs.keys():
                setattr(self, key, kwargs[key])
            # the reference overwrites the constructor's reference_id with the keyword default (:144-147)
        else:
            warnings.warn("in progress", UserWarning)

    # -- device plumbing -----------------------------------------------------------------------------
    def _state_shape(self, name):
        N, C, F, T, K = self.n_sources, self.n_channels, self.n_bins, self.n_frames, self.n_basis
        return {'basis': (F, K), 'activation': (K, T), 'latent': (N, K), 'spatial': (F, N, C, C), 'estimation': (N, F, T)}[name]

    def _state_dtype(self, name):
        return np.complex128 if name in ('spatial', 'estimation') else np.float64

    def _config(self):
        return dict(method=_lib.IS_MNMF, normalize=_lib.NORMALIZE_POWER if self.normalize else _lib.NORMALIZE_NONE, n_batch=1,
                    n_channels=self.n_channels, n_sources=self.n_sources, n_bins=self.n_bins, n_frames=self.n_frames,
                    n_basis=self.n_basis, reference_id=self.reference_id, eps=float(self.eps))

    def _prepare(self):
        X = self.input
        assert X is not None, "Specify data!"
        if self.author.lower() != 'sawada':
            raise NotImplementedError("Not support {}'s MNMF.".format(self.author))
        cfg = self._config()
        self._open_handle(tuple(sorted(cfg.items())), **cfg)
        self._send_input(X)
        self._push()

    # -- reference surface -----------------------------------------------------------------------------
    def _reset(self, **kwargs):
        super()._reset(**kwargs)

        author = self.author

        if author.lower() == 'sawada':
            self._reset_sawada()
        elif author.lower() == 'ozerov':
            raise NotImplementedError("Not support {}'s MNMF.".format(self.author))
        else:
            raise ValueError("Not support")

        self.__dict__['_input_token'] = None   # every __call__ re-reads the mixture, like the reference
        self._prepare()
        self._on_device.update(('basis', 'activation', 'latent', 'spatial', 'estimation'))
        self._device_changed('estimation')

    def _reset_sawada(self):
        """src/bss/mnmf.py:202-237: latent, basis, activation are drawn from the global NumPy state in that order."""
        n_basis = self.n_basis
        n_sources = self.n_sources
        eps = self.eps

        n_channels, n_bins, n_frames = self.input.shape

        if not hasattr(self, 'latent'):
            variance_latent = 1e-2
            Z = np.random.rand(n_sources, n_basis) * variance_latent + 1 / n_sources
            Zsum = Z.sum(axis=0)
            Zsum[Zsum < eps] = eps
            self.latent = Z / Zsum
        else:
            self.latent = np.array(self.latent, dtype=np.float64, copy=True)
        if not hasattr(self, 'spatial'):
            H = np.eye(n_channels)
            self.spatial = np.tile(H, reps=(n_bins, n_sources, 1, 1))
        else:
            self.spatial = np.array(self.spatial, copy=True)
        if not hasattr(self, 'basis'):
            self.basis = np.random.rand(n_bins, n_basis)
        else:
            self.basis = np.array(self.basis, dtype=np.float64, copy=True)
        if not hasattr(self, 'activation'):
            self.activation = np.random.rand(n_basis, n_frames)
        else:
            self.activation = np.array(self.activation, dtype=np.float64, copy=True)

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

        if self.callbacks is not None:
            for callback in self.callbacks:
                callback(self)

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

                for callback in self.callbacks:
                    callback(self)

        output = self.separate(self.input)
        self.estimation = output

        return output

    def __repr__(self):
        s = "IS-MNMF("
        s += "n_basis={n_basis}"
        if hasattr(self, 'n_sources'):
            s += ", n_sources={n_sources}"
        if hasattr(self, 'n_channels'):
            s += ", n_channels={n_channels}"
        s += ", normalize={normalize}"
        s += ", author={author}"
        s += ")"

        return s.format(**self.__dict__)

    def update_once(self):
        """update_basis_sawada, update_activation_sawada, update_latent_sawada, update_spatial_sawada (:311-315); the
        estimate the reference recomputes here (:308-309) is produced when `estimation` is read."""
        self._prepare()
        self._handle.update_once()
        self._device_changed()

    def reconstruct_covariance(self):
        """X_hat (n_bins, n_frames, n_channels, n_channels) of the current model (src/bss/mnmf.py:554-562); host-side
        convenience for callbacks, the device never stores it."""
        H, Z, T, V = self.spatial, self.latent, self.basis, self.activation
        HZ = np.einsum('fnij,nk->fkij', H, Z)
        return np.einsum('fkij,fk,kt->ftij', HZ, T, V)

    def compute_negative_loglikelihood(self):
        self._prepare()
        return float(self._handle.loss()[0])

    def separate(self, input):
        """Multichannel Wiener filter with the current model, image at `reference_id` (src/bss/mnmf.py:609-634).  `input` may
        be any mixture of the model's shape: a foreign one is put on the device for this call only."""
        return _separate_with_model(self, input)


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
        self._send_input(self.input, force=True)
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
        """Multichannel Wiener filter with the current model (src/bss/mnmf.py:919-946).  `input` may be any mixture of the
        model's shape: a foreign one is put on the device for this call only."""
        return _separate_with_model(self, input)


def _separate_with_model(model, input):
    own = model.input
    if own is None:
        raise AssertionError("Specify data!")
    input = np.asarray(input)
    if input.shape != np.shape(own):
        raise ValueError("input has shape {}, the model was fitted to {}".format(input.shape, np.shape(own)))
    model._prepare()
    shape = (model.n_sources, model.n_bins, model.n_frames)
    if input is own:
        return model._handle.separate(shape, np.complex128, projection_back=False)
    try:
        model._handle.set_input(input)
        return model._handle.separate(shape, np.complex128, projection_back=False)
    finally:
        model.__dict__['_input_token'] = None   # the model's own mixture goes back up before the next device operation


FastMNMF = FastMultichannelISNMF
MNMF = MultichannelISNMF
