"""Independent deeply learned matrix analysis on the GPU with the reference's class surface (src/sss/idlma.py).

`GaussIDLMA` (src/sss/idlma.py:88-258) alternates a source-model step -- a DNN supplied by the caller maps the power
spectrograms of the current estimates to source variances -- with the spatial-model step that every determined method of
the reference shares: weighted covariances `U[n,f] = mean_t x x^H / R[n,f,t]` followed by the gated IP sweep
(:175-210, the same code as src/bss/ilrma.py:497-530).  The DNN stays with the caller (a `torch.nn.Module`, exactly as
in the reference, or any callable on NumPy arrays); everything else -- covariance accumulate, IP sweep, projection-back
normalisation, negative log-likelihood, separation -- runs on the device behind the C ABI (method BSS_GAUSS_IDLMA): the
mixture stays resident, per iteration the estimates go to the host for the DNN and its variances come back.
"""
import numpy as np

from .. import _lib
from .._model import DeviceModel

EPS = 1e-12
THRESHOLD = 1e+12


class IDLMAbase(DeviceModel):
    """src/sss/idlma.py:10-86"""

    _STATE_IDS = {'demix_filter': _lib.STATE_DEMIX_FILTER, 'estimation': _lib.STATE_ESTIMATION}

    def __init__(self, normalize=True, callback=None, dnn_flooring=1e-5, eps=EPS):
        DeviceModel.__init__(self)
        self.callback = callback
        self.eps = eps
        self.input = None
        self.loss = []

        self.normalize = normalize
        self.dnn_flooring = dnn_flooring

    def _state_shape(self, name):
        if name == 'demix_filter':
            return (self.n_bins, self.n_sources, self.n_channels)
        if name == 'estimation':
            return (self.n_sources, self.n_bins, self.n_frames)
        raise KeyError(name)

    def _config(self):
        return dict(method=_lib.GAUSS_IDLMA, spatial=_lib.SPATIAL_IP,
                    normalize=_lib.NORMALIZE_PROJECTION_BACK if self.normalize == 'projection-back' else _lib.NORMALIZE_NONE,
                    n_batch=1, n_channels=self.n_channels, n_sources=self.n_sources, n_bins=self.n_bins,
                    n_frames=self.n_frames, n_basis=1, reference_id=getattr(self, 'reference_id', 0),
                    domain=float(getattr(self, 'domain', 2)), eps=float(self.eps),
                    threshold=float(getattr(self, 'threshold', THRESHOLD)))

    def _prepare(self):
        X = self.input
        assert X is not None, "Specify data!"
        cfg = self._config()
        if self._open_handle(tuple(sorted(cfg.items())), **cfg):
            self.__dict__['_variance_token'] = None
        if self._send_input(X):
            self.__dict__['_variance_token'] = None
        self._push()

    def _reset(self, dnn=None, **kwargs):
        assert self.input is not None, "Specify data!"

        for key in kwargs.keys():
            setattr(self, key, kwargs[key])

        X = self.input

        n_channels, n_bins, n_frames = X.shape
        n_sources = n_channels  # n_channels == n_sources

        self.n_sources, self.n_channels = n_sources, n_channels
        self.n_bins, self.n_frames = n_bins, n_frames

        # the reference re-creates W = I unconditionally (src/sss/idlma.py:34-36): presets are not honoured
        self._host.pop('demix_filter', None)
        self._dirty.discard('demix_filter')
        self._prepare()
        self._handle.reset_spatial()
        self._on_device.update(('demix_filter', 'estimation'))
        self._device_changed('demix_filter', 'estimation')

        self.dnn = dnn
        self.dnn_output = np.ones((n_sources, n_bins, n_frames))

    def __call__(self, input, iteration=100, **kwargs):
        """
        Args:
            input (n_channels, n_bins, n_frames)
        Returns:
            output (n_channels, n_bins, n_frames)
        """
        self.input = input

        self._reset(**kwargs)

        loss = self.compute_negative_loglikelihood()
        self.loss.append(loss)

        for idx in range(iteration):
            self.update_once()

            loss = self.compute_negative_loglikelihood()
            self.loss.append(loss)

            if self.callback is not None:
                self.callback(self)

        self._push()
        output = self._handle.separate((self.n_sources, self.n_bins, self.n_frames), np.complex128, projection_back=False)

        return output

    def update_once(self):
        raise NotImplementedError("Implement 'update_once' function")

    def separate(self, input, demix_filter):
        """
        Args:
            input (n_channels, n_bins, n_frames):
            demix_filter (n_bins, n_sources, n_channels):
        Returns:
            output (n_channels, n_bins, n_frames):
        """
        demix_filter = np.asarray(demix_filter)
        if demix_filter.ndim == 2:   # the reference's _reset passes one (N,C) matrix for all bins (src/sss/idlma.py:34-36)
            demix_filter = np.tile(demix_filter, reps=(input.shape[1], 1, 1))
        return _lib.demix(input, demix_filter)

    def compute_negative_loglikelihood(self):
        raise NotImplementedError("Implement 'compute_negative_loglikelihood' function.")


class GaussIDLMA(IDLMAbase):
    """Drop-in for src/sss/idlma.py:88-258."""

    def __init__(self, domain=2, normalize='power', reference_id=0, callback=None, dnn_flooring=1e-5, eps=EPS, threshold=THRESHOLD):
        """
        Args:
            normalize <str>: 'power': power based normalization, or 'projection-back': projection back based normalization.
            threshold <float>: threshold for condition number when computing (WU)^{-1}.
        """
        super().__init__(normalize=normalize, callback=callback, dnn_flooring=dnn_flooring, eps=eps)

        assert 1 <= domain <= 2, "1 <= `domain` <= 2 is not satisfied."

        self.domain = domain
        self.reference_id = reference_id
        self.threshold = threshold

    def __call__(self, input, iteration=100, **kwargs):
        """
        Args:
            input (n_channels, n_bins, n_frames)
        Returns:
            output (n_channels, n_bins, n_frames)
        """
        self.input = input

        self._reset(**kwargs)

        loss = self.compute_negative_loglikelihood()
        self.loss.append(loss)

        for idx in range(iteration):
            self.update_once()

            loss = self.compute_negative_loglikelihood()
            self.loss.append(loss)

            if self.callback is not None:
                self.callback(self)

        self._push()
        output = self._handle.separate((self.n_sources, self.n_bins, self.n_frames), np.complex128, projection_back=True)
        self._host['estimation'] = output

        return output

    def update_once(self, is_source_model_update=True):
        if is_source_model_update:
            self.update_source_model()
        self.update_space_model()

        # src/sss/idlma.py:150-162: only 'projection-back' is implemented upstream; the checks come after the sweep
        if self.normalize:
            if self.normalize == 'projection-back':
                self._handle.normalize()
                self._device_changed()
            else:
                raise ValueError("Not support normalization based on {}. Choose 'power' or 'projection-back'".format(self.normalize))
        else:
            raise ValueError("Set normalize=True")

    def update_source_model(self):
        """src/sss/idlma.py:167-173: power of the current estimates -> DNN -> (floored) variances."""
        self._prepare()
        Y = self._handle.separate((self.n_sources, self.n_bins, self.n_frames), np.complex128, projection_back=False)
        P = np.abs(Y)**2

        dnn_output = self.estimate_by_dnn(P)
        self.dnn_output = dnn_output

        if self.dnn_flooring:
            self.floor_dnn_output()

    def update_space_model(self):
        """src/sss/idlma.py:175-210 on the device."""
        self._prepare()
        self._send_variance()
        self._handle.update_once()
        self._device_changed()

    def _send_variance(self):
        """R = dnn_output^(2/domain) (src/sss/idlma.py:181, :253), uploaded when it changed; the eps floor is applied on the device."""
        d = np.asarray(self.dnn_output)
        token = (id(self.dnn_output), d.shape, float(d.sum()), float(self.domain))
        if self.__dict__.get('_variance_token') == token:
            return
        if d.shape != (self.n_sources, self.n_bins, self.n_frames):
            raise ValueError("dnn_output has shape {}, expected {}".format(d.shape, (self.n_sources, self.n_bins, self.n_frames)))
        R = d**(2 / self.domain)   # in the array's own precision, like the reference (float32 for a torch DNN)
        self._handle.set_state(_lib.STATE_VARIANCE, np.ascontiguousarray(R, dtype=np.float64), np.float64)
        self.__dict__['_variance_token'] = token

    def estimate_by_dnn(self, input):
        """src/sss/idlma.py:212-226.  `self.dnn` is a torch.nn.Module as in the reference; a plain callable on NumPy arrays is
        accepted as well."""
        domain = self.domain
        input = input**(domain / 2)

        dnn = self.dnn
        if hasattr(dnn, 'parameters'):
            import torch
            with torch.no_grad():
                input = torch.Tensor(input)
                if next(dnn.parameters()).is_cuda:
                    input = input.cuda()
                output = dnn(input)
            output = output.cpu().numpy()
        else:
            output = np.asarray(dnn(input))
        output = output**(2 / domain)

        return output

    def floor_dnn_output(self):
        floor = self.dnn_flooring
        dnn_output = self.dnn_output
        dnn_output = np.maximum(dnn_output, floor)
        self.dnn_output = dnn_output

    def compute_negative_loglikelihood(self):
        """src/sss/idlma.py:244-258, reduced on the device."""
        self._prepare()
        self._send_variance()
        return float(self._handle.loss()[0])
