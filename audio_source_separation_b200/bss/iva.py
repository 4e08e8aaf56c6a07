"""Auxiliary-function IVA on the GPU with the reference's class surface (src/bss/iva.py).

`AuxLaplaceIVA` (src/bss/iva.py:388-619) and `AuxGaussIVA` (:621-802), spatial updates 'IP'/'IP1', 'ISS'
and 'IP2'/'pairwise'.  Gradient and proximal IVA (src/bss/iva.py:130-287, :831-951) are not
auxiliary-function updates and are outside this package.
"""
import numpy as np

from .. import _lib
from .._model import DeviceModel, parse_spatial

EPS = 1e-12
THRESHOLD = 1e+12

__algorithms_spatial__ = ['IP', 'IVA', 'ISS', 'IPA', 'pairwise', 'IP1', 'IP2']


class IVAbase(DeviceModel):
    """src/bss/iva.py:22-128"""

    _STATE_IDS = {'demix_filter': _lib.STATE_DEMIX_FILTER, 'estimation': _lib.STATE_ESTIMATION}
    _method = None

    def __init__(self, callbacks=None, recordable_loss=True, eps=EPS):
        DeviceModel.__init__(self)
        if callbacks is not None:
            if callable(callbacks):
                callbacks = [callbacks]
            self.callbacks = callbacks
        else:
            self.callbacks = None
        self.eps = eps

        self.input = None
        self.recordable_loss = recordable_loss
        if self.recordable_loss:
            self.loss = []
        else:
            self.loss = None

    def _state_shape(self, name):
        if name == 'demix_filter':
            return (self.n_bins, self.n_sources, self.n_channels)
        if name == 'estimation':
            return (self.n_sources, self.n_bins, self.n_frames)
        raise KeyError(name)

    def _config(self):
        return dict(method=self._method, spatial=parse_spatial(getattr(self, 'algorithm_spatial', 'IP')),
                    normalize=_lib.NORMALIZE_NONE, n_batch=1, n_channels=self.n_channels, n_sources=self.n_sources,
                    n_bins=self.n_bins, n_frames=self.n_frames, n_basis=1, reference_id=getattr(self, 'reference_id', 0),
                    eps=float(self.eps), threshold=float(getattr(self, 'threshold', THRESHOLD)))

    def _prepare(self):
        X = self.input
        assert X is not None, "Specify data!"
        cfg = self._config()
        self._open_handle(tuple(sorted(cfg.items())), **cfg)
        self._send_input(X)
        self._push()

    def _reset(self, **kwargs):
        assert self.input is not None, "Specify data!"

        for key in kwargs.keys():
            setattr(self, key, kwargs[key])

        X = self.input

        n_channels, n_bins, n_frames = X.shape
        n_sources = n_channels  # n_channels == n_sources

        self.n_sources, self.n_channels = n_sources, n_channels
        self.n_bins, self.n_frames = n_bins, n_frames

        preset_filter = hasattr(self, 'demix_filter') and self.demix_filter is not None
        if preset_filter:
            self.demix_filter = np.array(self.demix_filter, dtype=np.complex128, copy=True)

        cfg = self._config()
        self._open_handle(tuple(sorted(cfg.items())), **cfg)
        self._send_input(X, force=True)
        if not preset_filter:
            self._handle.reset_spatial()
            self._host.pop('demix_filter', None)
            self._dirty.discard('demix_filter')
        self._push()
        self._on_device.update(('demix_filter', 'estimation'))
        self._device_changed('estimation')

    def separate(self, input, demix_filter):
        """
        Args:
            input (n_channels, n_bins, n_frames):
            demix_filter (n_bins, n_sources, n_channels):
        Returns:
            output (n_channels, n_bins, n_frames):
        """
        return _lib.demix(input, demix_filter)

    def compute_demix_filter(self, estimation, input):
        """W = Y X^H (X X^H)^-1 per bin (src/bss/iva.py:119-125); served from the device state when the arguments are the model's own
        estimation / input, from one least-squares kernel over the two arrays otherwise."""
        own = self._handle is not None and input is self.input and 'estimation' not in self._dirty and (
            estimation is self._host.get('estimation') or estimation is None)
        if own and self.algorithm_spatial == 'ISS':
            self._push()
            self._handle.compute_demix_filter()
            return self._handle.get_state(_lib.STATE_DEMIX_FILTER, self._state_shape('demix_filter'), np.complex128)
        return np.ascontiguousarray(_lib.least_squares_map(estimation, input).transpose(2, 0, 1))

    def update_once(self):
        raise NotImplementedError("Implement 'update_once' function.")

    def compute_negative_loglikelihood(self):
        raise NotImplementedError("Implement 'compute_negative_loglikelihood' function.")


class AuxIVAbase(IVAbase):
    """src/bss/iva.py:289-386"""

    def __init__(self, algorithm_spatial='IP', reference_id=0, callbacks=None, apply_projection_back=True, recordable_loss=True,
                 eps=EPS, threshold=THRESHOLD):
        super().__init__(callbacks=callbacks, recordable_loss=recordable_loss, eps=eps)

        self.algorithm_spatial = algorithm_spatial
        self.reference_id = reference_id
        self.apply_projection_back = apply_projection_back
        self.threshold = threshold

        if self.algorithm_spatial not in __algorithms_spatial__:
            raise ValueError("Not support {} based spatial updates.".format(self.algorithm_spatial))

        if self.algorithm_spatial in ['pairwise', 'IP2']:
            self.update_pair = None

    def _reset(self, **kwargs):
        super()._reset(**kwargs)

        if self.algorithm_spatial == 'ISS':
            self._host['demix_filter'] = None

    def _run_callbacks(self):
        if self.callbacks is None:
            return
        if self.algorithm_spatial == 'ISS':
            self._handle.compute_demix_filter()
            self._host.pop('demix_filter', None)
        for callback in self.callbacks:
            callback(self)
        if self.algorithm_spatial == 'ISS':
            self._push()
            self._host['demix_filter'] = None

    def __call__(self, input, iteration=100, **kwargs):
        """
        Args:
            input (n_channels, n_bins, n_frames)
        Returns:
            output (n_channels, n_bins, n_frames)
        """
        self.input = input

        self._reset(**kwargs)

        if self.recordable_loss:
            loss = self.compute_negative_loglikelihood()
            self.loss.append(loss)

        self._run_callbacks()

        if self.callbacks is None:
            self._check_spatial()
            self._push()
            if self.algorithm_spatial in ['pairwise', 'IP2']:
                # also when it is None: a reused handle must not continue an earlier run's schedule
                self._handle.set_update_pair(*(self.update_pair if self.update_pair is not None else (-1, -1)))
            if self.recordable_loss:
                # the loss after every iteration is reduced on the device and fetched once
                self.loss.extend(float(v) for v in self._handle.run_record(iteration)[:, 0])
            else:
                self._handle.run(iteration)
            if self.algorithm_spatial in ['pairwise', 'IP2']:
                for _ in range(iteration):
                    self._select_update_pair(tell_device=False)
            self._after_update()
        else:
            for idx in range(iteration):
                if self.algorithm_spatial in ['pairwise', 'IP2']:
                    self._select_update_pair()

                self.update_once()

                if self.recordable_loss:
                    loss = self.compute_negative_loglikelihood()
                    self.loss.append(loss)

                self._run_callbacks()

        self._push()
        if self.algorithm_spatial == 'ISS':
            self._handle.compute_demix_filter()
            self._host.pop('demix_filter', None)
        output = self._handle.separate((self.n_sources, self.n_bins, self.n_frames), np.complex128,
                                       projection_back=bool(self.apply_projection_back))
        self._host['estimation'] = output

        return output

    def __repr__(self):
        s = "AuxIVA("
        s += "algorithm_spatial={algorithm_spatial}"
        s += ")"

        return s.format(**self.__dict__)

    def _after_update(self):
        self._device_changed()
        if self.algorithm_spatial == 'ISS':
            self._host['demix_filter'] = None

    def _check_spatial(self):
        if self.algorithm_spatial == 'IPA':
            raise NotImplementedError("In progress...")
        if self.algorithm_spatial not in ['IP', 'IP1', 'ISS', 'pairwise', 'IP2']:
            raise ValueError("Not support {} based spatial updates.".format(self.algorithm_spatial))

    def update_once(self):
        self._check_spatial()
        self._prepare()
        self._handle.update_once()
        self._after_update()

    def _select_update_pair(self, tell_device=True):
        """src/bss/iva.py:372-383"""
        n_sources = self.n_sources

        if self.update_pair is None:
            m, n = 0, 1
        else:
            m, n = self.update_pair
            m, n = m + 1, n + 1
            m, n = m % n_sources, n % n_sources

        self.update_pair = m, n
        if tell_device and self._handle is not None:
            self._handle.set_update_pair(m, n)

    def compute_negative_loglikelihood(self):
        self._prepare()
        return float(self._handle.loss()[0])


class AuxLaplaceIVA(AuxIVAbase):
    """Drop-in for src/bss/iva.py:388-619."""
    _method = _lib.AUX_LAPLACE_IVA

    def __init__(self, algorithm_spatial='IP', reference_id=0, callbacks=None, apply_projection_back=True, recordable_loss=True,
                 eps=EPS, threshold=THRESHOLD):
        super().__init__(algorithm_spatial=algorithm_spatial, reference_id=reference_id, callbacks=callbacks,
                         apply_projection_back=apply_projection_back, recordable_loss=recordable_loss, eps=eps,
                         threshold=threshold)

    def __repr__(self):
        s = "AuxLaplaceIVA("
        s += "algorithm_spatial={algorithm_spatial}"
        s += ")"

        return s.format(**self.__dict__)


class AuxGaussIVA(AuxIVAbase):
    """Drop-in for src/bss/iva.py:621-802."""
    _method = _lib.AUX_GAUSS_IVA

    def __init__(self, algorithm_spatial='IP', reference_id=0, callbacks=None, apply_projection_back=True, recordable_loss=True,
                 eps=EPS, threshold=THRESHOLD):
        super().__init__(algorithm_spatial=algorithm_spatial, reference_id=reference_id, callbacks=callbacks,
                         apply_projection_back=apply_projection_back, recordable_loss=recordable_loss, eps=eps,
                         threshold=threshold)

    def _check_spatial(self):
        if self.algorithm_spatial in ['pairwise', 'IP2']:
            raise NotImplementedError("In progress...")   # src/bss/iva.py:777-778
        super()._check_spatial()

    def __repr__(self):
        s = "AuxGaussIVA("
        s += "algorithm_spatial={algorithm_spatial}"
        s += ")"

        return s.format(**self.__dict__)
