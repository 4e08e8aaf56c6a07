"""ILRMA family on the GPU with the reference's class surface (src/bss/ilrma.py).

`GaussILRMA` (src/bss/ilrma.py:178) and `tILRMA` (:713) keep the reference's constructor arguments,
`__call__`, `update_once`, `separate`, `compute_negative_loglikelihood`, callbacks and public state
attributes; every update dispatches through libbssgpu's C ABI into CUDA kernels.  Random initial state
is drawn on the host from NumPy's global legacy RNG in the reference's order (:79-104) so that
`np.random.seed` reproduces the reference's starting point bit for bit.
"""
import warnings

import numpy as np

from .. import _lib
from .._model import DeviceModel, parse_spatial, parse_normalize

EPS = 1e-12
THRESHOLD = 1e+12

__algorithms_spatial__ = ['IP', 'IVA', 'ISS', 'IPA', 'pairwise', 'IP1', 'IP2']


class ILRMAbase(DeviceModel):
    """Independent low-rank matrix analysis (src/bss/ilrma.py:22-176)."""

    _STATE_IDS = {'demix_filter': _lib.STATE_DEMIX_FILTER, 'estimation': _lib.STATE_ESTIMATION,
                  'basis': _lib.STATE_BASIS, 'activation': _lib.STATE_ACTIVATION, 'latent': _lib.STATE_LATENT}
    _method = None

    def __init__(self, n_basis=10, partitioning=False, normalize=True, algorithm_spatial='IP', callbacks=None,
                 recordable_loss=True, eps=EPS):
        DeviceModel.__init__(self)
        if callbacks is not None:
            if callable(callbacks):
                callbacks = [callbacks]
            self.callbacks = callbacks
        else:
            self.callbacks = None
        self.eps = eps

        self.n_basis = n_basis
        self.partitioning = partitioning
        self.normalize = normalize

        assert algorithm_spatial in __algorithms_spatial__, "Choose from {} as `algorithm_spatial`.".format(__algorithms_spatial__)
        assert algorithm_spatial in ['IP', 'ISS', 'pairwise', 'IP1', 'IP2'], "Not support {}-based demixing filter updates.".format(algorithm_spatial)
        self.algorithm_spatial = algorithm_spatial

        self.input = None
        self.recordable_loss = recordable_loss
        if self.recordable_loss:
            self.loss = []
        else:
            self.loss = None

    # -- shapes ----------------------------------------------------------------------------------
    def _state_shape(self, name):
        N, C, F, T, K = self.n_sources, self.n_channels, self.n_bins, self.n_frames, self.n_basis
        if name == 'demix_filter':
            return (F, N, C)
        if name == 'estimation':
            return (N, F, T)
        if name == 'basis':
            return (F, K) if self.partitioning else (N, F, K)
        if name == 'activation':
            return (K, T) if self.partitioning else (N, K, T)
        if name == 'latent':
            return (N, K)
        raise KeyError(name)

    def _config(self):
        return dict(method=self._method, spatial=parse_spatial(self.algorithm_spatial),
                    normalize=parse_normalize(self.normalize), partitioning=1 if self.partitioning else 0,
                    n_batch=1, n_channels=self.n_channels, n_sources=self.n_sources, n_bins=self.n_bins,
                    n_frames=self.n_frames, n_basis=self.n_basis, reference_id=getattr(self, 'reference_id', 0),
                    domain=float(getattr(self, 'domain', 2)), nu=float(getattr(self, 'nu', 1)), eps=float(self.eps),
                    threshold=float(getattr(self, 'threshold', THRESHOLD)))

    def _prepare(self):
        """Make the device agree with the host attributes (handle, input, assigned state)."""
        X = self.input
        assert X is not None, "Specify data!"
        cfg = self._config()
        key = tuple(sorted(cfg.items()))
        self._open_handle(key, **cfg)
        self._send_input(X)
        self._push()

    # -- reference surface -----------------------------------------------------------------------
    def _reset(self, **kwargs):
        assert self.input is not None, "Specify data!"

        for key in kwargs.keys():
            setattr(self, key, kwargs[key])

        n_basis = self.n_basis
        eps = self.eps

        X = self.input

        n_channels, n_bins, n_frames = X.shape
        n_sources = n_channels  # n_channels == n_sources

        self.n_sources, self.n_channels = n_sources, n_channels
        self.n_bins, self.n_frames = n_bins, n_frames

        iss = self.algorithm_spatial == 'ISS'
        preset_filter = hasattr(self, 'demix_filter') and self.demix_filter is not None
        if preset_filter:
            self.demix_filter = np.array(self.demix_filter, dtype=np.complex128, copy=True)

        # random initial source model, drawn in the reference's order (latent, basis, activation)
        if self.partitioning:
            if not hasattr(self, 'latent'):
                variance_latent = 1e-2
                Z = np.random.rand(n_sources, n_basis) * variance_latent + 1 / n_sources
                Zsum = Z.sum(axis=0)
                Zsum[Zsum < eps] = eps
                self.latent = Z / Zsum
            else:
                self.latent = np.array(self.latent, dtype=np.float64, copy=True)
            if not hasattr(self, 'basis'):
                self.basis = np.random.rand(n_bins, n_basis)
            else:
                self.basis = np.array(self.basis, dtype=np.float64, copy=True)
            if not hasattr(self, 'activation'):
                self.activation = np.random.rand(n_basis, n_frames)
            else:
                self.activation = np.array(self.activation, dtype=np.float64, copy=True)
        else:
            if not hasattr(self, 'basis'):
                self.basis = np.random.rand(n_sources, n_bins, n_basis)
            else:
                self.basis = np.array(self.basis, dtype=np.float64, copy=True)
            if not hasattr(self, 'activation'):
                self.activation = np.random.rand(n_sources, n_basis, n_frames)
            else:
                self.activation = np.array(self.activation, dtype=np.float64, copy=True)

        # device side: handle, input, W (identity unless preset), estimation = separate(X, W)
        cfg = self._config()
        self._open_handle(tuple(sorted(cfg.items())), **cfg)
        self._send_input(X, force=True)
        if not preset_filter:
            self._handle.reset_spatial()
            self._host.pop('demix_filter', None)
            self._dirty.discard('demix_filter')
        self._push()
        self._on_device.update(('demix_filter', 'estimation', 'basis', 'activation'))
        if self.partitioning:
            self._on_device.add('latent')
        self._device_changed('estimation')
        if iss:
            self._host['demix_filter'] = None

    def _run_callbacks(self):
        if self.callbacks is None:
            return
        if self.algorithm_spatial == 'ISS':
            # the ISS update keeps no filter: rebuild it for the callbacks, as the reference does
            self._handle.compute_demix_filter()
            self._host.pop('demix_filter', None)
            for callback in self.callbacks:
                callback(self)
            self._push()
            self._host['demix_filter'] = None
        else:
            for callback in self.callbacks:
                callback(self)

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
            # no callback observes the intermediate states: run the whole loop on the device
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
        output = self._handle.separate((self.n_sources, self.n_bins, self.n_frames), np.complex128, projection_back=True)
        self._host['estimation'] = output

        return output

    def __repr__(self):
        s = "ILRMA("
        s += "n_basis={n_basis}"
        s += ", partitioning={partitioning}"
        s += ", normalize={normalize}"
        s += ")"

        return s.format(**self.__dict__)

    def _after_update(self):
        self._device_changed()
        if self.algorithm_spatial == 'ISS':
            self._host['demix_filter'] = None

    def update_once(self):
        if self.normalize and self.normalize not in ('power', 'projection-back'):
            raise ValueError("Not support normalization based on {}. Choose 'power' or 'projection-back'".format(self.normalize))
        self._prepare()
        self._handle.update_once()
        self._after_update()

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
        """W = Y X^H (X X^H)^-1 per bin (src/bss/ilrma.py:167-173); served from the device state when the arguments are the model's own
        estimation / input, from one least-squares kernel over the two arrays otherwise."""
        own = self._handle is not None and input is self.input and 'estimation' not in self._dirty and (
            estimation is self._host.get('estimation') or estimation is None)
        if own and self.algorithm_spatial == 'ISS':
            self._push()
            self._handle.compute_demix_filter()
            return self._handle.get_state(_lib.STATE_DEMIX_FILTER, self._state_shape('demix_filter'), np.complex128)
        return np.ascontiguousarray(_lib.least_squares_map(estimation, input).transpose(2, 0, 1))

    def compute_negative_loglikelihood(self):
        self._prepare()
        return float(self._handle.loss()[0])

    def _select_update_pair(self, tell_device=True):
        """src/bss/ilrma.py:635-646"""
        n_sources = self.n_sources
        if self.update_pair is None:
            m, n = 0, 1
        else:
            m, n = self.update_pair
            m, n = (m + 1) % n_sources, (n + 1) % n_sources
        self.update_pair = m, n
        if tell_device and self._handle is not None:
            self._handle.set_update_pair(m, n)


class GaussILRMA(ILRMAbase):
    """
    Reference: "Determined Blind Source Separation Unifying Independent Vector Analysis and Nonnegative Matrix Factorization"
    Drop-in for src/bss/ilrma.py:178-677.
    """
    _method = _lib.GAUSS_ILRMA

    def __init__(self, n_basis=10, domain=2, partitioning=False, normalize='power', algorithm_spatial='IP', reference_id=0,
                 callbacks=None, recordable_loss=True, eps=EPS, threshold=THRESHOLD):
        """
        Args:
            normalize <str>: 'power': power based normalization, or 'projection-back': projection back based normalization.
            threshold <float>: threshold for condition number when computing (WU)^{-1}.
        """
        super().__init__(n_basis=n_basis, partitioning=partitioning, normalize=normalize, algorithm_spatial=algorithm_spatial,
                         callbacks=callbacks, recordable_loss=recordable_loss, eps=eps)

        assert 1 <= domain <= 2, "1 <= `domain` <= 2 is not satisfied."

        self.domain = domain
        self.reference_id = reference_id
        self.threshold = threshold

        if self.algorithm_spatial == 'ISS':
            warnings.warn("in progress", UserWarning)

        if self.algorithm_spatial in ['pairwise', 'IP2']:
            self.update_pair = None

    def update_once(self):
        if self.normalize == 'projection-back' and self.partitioning:
            raise NotImplementedError("Not support 'projection-back' based normalization for partitioninig function. Choose 'power' based normalization.")
        if self.partitioning:
            assert self.domain == 2, "Not support domain = {}".format(self.domain)
        super().update_once()

    def __repr__(self):
        s = "Gauss-ILRMA("
        s += "n_basis={n_basis}"
        s += ", domain={domain}"
        s += ", partitioning={partitioning}"
        s += ", normalize={normalize}"
        s += ", algorithm_spatial={algorithm_spatial}"
        s += ")"

        return s.format(**self.__dict__)


class ConsistentGaussILRMA(GaussILRMA):
    """
    Reference: "Consistent independent low-rank matrix analysis for determined blind source separation"
    Drop-in for src/bss/ilrma.py:1102-1233.  The reference's update_once first replaces `estimation` by its STFT-consistent
    projection (:1206-1207), but its IP path then recomputes the estimates from `demix_filter` (:360-364), so the projection
    never reaches the update ("reference quirks"): what remains is the source model, the IP sweep and
    a projection-back rescaling of W and T with exponent 2 every iteration (:1219-1233) -- which is what runs here.
    `fft_size` / `hop_size` are kept as attributes; `transform.stft` provides the GPU STFT/ISTFT pair.
    """

    def __init__(self, n_basis=10, partitioning=False, algorithm_spatial='IP', reference_id=0, fft_size=None, hop_size=None,
                 callbacks=None, recordable_loss=True, eps=EPS, threshold=THRESHOLD):
        super().__init__(n_basis=n_basis, partitioning=partitioning, normalize=False, algorithm_spatial=algorithm_spatial,
                         reference_id=reference_id, callbacks=callbacks, recordable_loss=recordable_loss, eps=eps,
                         threshold=threshold)

        if fft_size is None:
            raise ValueError("Specify `fft_size`.")

        if hop_size is None:
            hop_size = fft_size // 2

        self.fft_size, self.hop_size = fft_size, hop_size

        assert self.algorithm_spatial == 'IP', "Supports only IP-based spatial update."

    def _config(self):
        cfg = super()._config()
        cfg['normalize'] = _lib.NORMALIZE_PROJECTION_BACK   # src/bss/ilrma.py:1219-1226, unconditional
        cfg['domain'] = 2.0
        return cfg

    def update_once(self):
        if self.partitioning:
            raise NotImplementedError("Not support 'projection-back' based normalization for partitioninig function. Choose 'power' based normalization.")
        ILRMAbase.update_once(self)

    def __repr__(self):
        s = "Consistent-GaussILRMA("
        s += "n_basis={n_basis}"
        s += ", domain={domain}"
        s += ", partitioning={partitioning}"
        s += ", normalize={normalize}"
        s += ", algorithm_spatial={algorithm_spatial}"
        s += ")"

        return s.format(**self.__dict__)


class tILRMA(ILRMAbase):
    """
    Reference: "Independent low-rank matrix analysis based on complex student's t-distribution for blind audio source separation"
    Drop-in for src/bss/ilrma.py:713-1020.
    """
    _method = _lib.T_ILRMA

    def __init__(self, n_basis=10, nu=1, domain=2, partitioning=False, normalize='power', algorithm_spatial='IP', reference_id=0,
                 callbacks=None, recordable_loss=True, eps=EPS):
        """
        Args:
            nu: degree of freedom. nu = 1: Cauchy distribution, nu -> infty: Gaussian distribution.
        """
        super().__init__(n_basis=n_basis, partitioning=partitioning, normalize=normalize, algorithm_spatial=algorithm_spatial,
                         callbacks=callbacks, recordable_loss=recordable_loss, eps=eps)

        self.nu = nu
        self.domain = domain
        self.reference_id = reference_id

        assert self.algorithm_spatial == 'IP', "Supports only IP-based spatial update."

    def update_once(self):
        if self.normalize and self.normalize != 'power':
            raise ValueError("Not support normalization based on {}. Choose 'power' or 'projection-back'".format(self.normalize))
        assert self.domain == 2, "Only 'domain' = 2 is supported."
        super().update_once()

    def __repr__(self):
        s = "t-ILRMA("
        s += "n_basis={n_basis}"
        s += ", nu={nu}"
        s += ", domain={domain}"
        s += ", partitioning={partitioning}"
        s += ", normalize={normalize}"
        s += ", algorithm_spatial={algorithm_spatial}"
        s += ")"

        return s.format(**self.__dict__)


class GGDILRMA(ILRMAbase):
    """src/bss/ilrma.py:679-711: not implemented upstream either."""

    def __init__(self, n_basis=10, beta=1, domain=2, partitioning=False, normalize='power', algorithm_spatial='IP', reference_id=0,
                 callbacks=None, recordable_loss=True, eps=EPS):
        super().__init__(n_basis=n_basis, partitioning=partitioning, normalize=normalize, algorithm_spatial=algorithm_spatial,
                         callbacks=callbacks, recordable_loss=recordable_loss, eps=eps)
        raise NotImplementedError("In progress")


class KLILRMA(ILRMAbase):
    """src/bss/ilrma.py:1022-1082: not implemented upstream either."""

    def __init__(self, n_basis=10, partitioning=False, normalize='power', algorithm_spatial='IP', reference_id=0, callbacks=None,
                 recordable_loss=True, eps=EPS):
        super().__init__(n_basis=n_basis, partitioning=partitioning, normalize=normalize, algorithm_spatial=algorithm_spatial,
                         callbacks=callbacks, recordable_loss=recordable_loss, eps=eps)
        raise NotImplementedError("In progress")


class RegularizedILRMA(ILRMAbase):
    """src/bss/ilrma.py:1084-1100: not implemented upstream either."""

    def __init__(self, n_basis=10, partitioning=False, normalize='power', algorithm_spatial='IP', reference_id=0, callbacks=None,
                 recordable_loss=True, eps=EPS):
        super().__init__(n_basis=n_basis, partitioning=partitioning, normalize=normalize, algorithm_spatial=algorithm_spatial,
                         callbacks=callbacks, recordable_loss=recordable_loss, eps=eps)
        raise NotImplementedError("In progress")


def update_spatial_model_ip(input, demix_filter, variance, threshold=THRESHOLD, eps=EPS):
    """One iterative-projection sweep with externally supplied source variances, the spatial half of every determined
    method of the reference: `GaussILRMA.update_spatial_model_ip` (src/bss/ilrma.py:483-535) once R = (T V)^(2/domain) is
    known, and verbatim `GaussIDLMA.update_space_model` (src/sss/idlma.py:175-210), whose R comes from a DNN.

    Args:
        input (n_channels, n_bins, n_frames): mixture
        demix_filter (n_bins, n_sources, n_channels): current filters
        variance (n_sources, n_bins, n_frames): source variances R (floored at `eps` here, as the reference does)
    Returns:
        demix_filter (n_bins, n_sources, n_channels), estimation (n_sources, n_bins, n_frames)
    """
    R = np.array(variance, dtype=np.float64, copy=True)
    R[R < eps] = eps
    U = _lib.weighted_covariance(input, R)
    W, _ = _lib.ip_update(demix_filter, U, threshold=threshold)
    return W, _lib.demix(input, W)
