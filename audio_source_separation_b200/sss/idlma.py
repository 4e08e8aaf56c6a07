"""Independent deeply learned matrix analysis on the GPU with the reference's class surface (src/sss/idlma.py).

`GaussIDLMA` (src/sss/idlma.py:88-258) alternates a source-model step -- a DNN supplied by the caller maps the power
spectrograms of the current estimates to source variances -- with the spatial-model step that every determined method of
the reference shares: weighted covariances `U[n,f] = mean_t x x^H / R[n,f,t]` followed by the gated IP sweep
(:175-210, the same code as src/bss/ilrma.py:497-530).  The DNN stays with the caller (a `torch.nn.Module`, exactly as
in the reference, or any callable on NumPy arrays); everything else -- covariance accumulate, IP sweep, projection-back
normalisation, negative log-likelihood, separation -- runs on the device behind the C ABI (method BSS_GAUSS_IDLMA): the
mixture stays resident, per iteration the estimates go to the host for the DNN and its variances come back.

Same constructor arguments, attributes (`demix_filter`, `estimation`, `dnn`, `dnn_output`, `loss`, `input`), methods and
exceptions as the reference, including its quirks: `_reset` always restarts from W = I (:34-36), `loss` keeps growing
across calls (:15), and only normalize='projection-back' survives `update_once` -- the default 'power' raises after the
spatial update has been applied (:159-162).
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
        self.normalize, self.dnn_flooring = normalize, dnn_flooring
        self.callback, self.eps = callback, eps
        self.input, self.loss = None, []

    # -- device plumbing -------------------------------------------------------------------------------
    def _state_shape(self, name):
        shapes = {'demix_filter': (self.n_bins, self.n_sources, self.n_channels),
                  'estimation': (self.n_sources, self.n_bins, self.n_frames)}
        return shapes[name]

    def _config(self):
        pb = self.normalize == 'projection-back'
        return dict(method=_lib.GAUSS_IDLMA, spatial=_lib.SPATIAL_IP,
                    normalize=_lib.NORMALIZE_PROJECTION_BACK if pb else _lib.NORMALIZE_NONE,
                    n_batch=1, n_channels=self.n_channels, n_sources=self.n_sources, n_bins=self.n_bins,
                    n_frames=self.n_frames, n_basis=1, reference_id=getattr(self, 'reference_id', 0),
                    domain=float(getattr(self, 'domain', 2)), eps=float(self.eps),
                    threshold=float(getattr(self, 'threshold', THRESHOLD)))

    def _prepare(self):
        """Handle for the current shape / configuration, mixture resident, pending host edits uploaded."""
        if self.input is None:
            raise AssertionError("Specify data!")
        cfg = self._config()
        fresh = self._open_handle(tuple(sorted(cfg.items())), **cfg)
        fresh = self._send_input(self.input) or fresh
        if fresh:
            self.__dict__['_variance_token'] = None   # a new handle / mixture has no variances yet
        self._push()

    def _estimates(self, projection_back):
        return self._handle.separate((self.n_sources, self.n_bins, self.n_frames), np.complex128, projection_back=projection_back)

    def _iterate(self, input, iteration, kwargs):
        """The loop both `__call__`s of the reference run (src/sss/idlma.py:41-67 and :105-128)."""
        self.input = input
        self._reset(**kwargs)
        self.loss.append(self.compute_negative_loglikelihood())
        for _ in range(iteration):
            self.update_once()
            self.loss.append(self.compute_negative_loglikelihood())
            if self.callback is not None:
                self.callback(self)
        self._push()

    # -- reference surface ---------------------------------------------------------------------------------
    def _reset(self, dnn=None, **kwargs):
        assert self.input is not None, "Specify data!"
        for key, value in kwargs.items():
            setattr(self, key, value)
        self.n_channels, self.n_bins, self.n_frames = self.input.shape
        self.n_sources = self.n_channels   # determined case only
        # W = I for every bin whatever was preset, estimation = separate(X, W) (src/sss/idlma.py:34-36)
        self._host.pop('demix_filter', None)
        self._dirty.discard('demix_filter')
        self.__dict__['_input_token'] = None   # every __call__ re-reads the mixture, like the reference
        self._prepare()
        self._handle.reset_spatial()
        self._on_device.update(('demix_filter', 'estimation'))
        self._device_changed('demix_filter', 'estimation')
        self.dnn = dnn
        self.dnn_output = np.ones((self.n_sources, self.n_bins, self.n_frames))

    def __call__(self, input, iteration=100, **kwargs):
        """input (n_channels, n_bins, n_frames) -> separate(input, demix_filter) after `iteration` updates."""
        self._iterate(input, iteration, kwargs)
        return self._estimates(projection_back=False)

    def update_once(self):
        raise NotImplementedError("Implement 'update_once' function")

    def separate(self, input, demix_filter):
        """input (n_channels, n_bins, n_frames), demix_filter (n_bins, n_sources, n_channels) or one (n_sources, n_channels)
        matrix for all bins, as the reference's `_reset` passes it -> (n_sources, n_bins, n_frames)."""
        demix_filter = np.asarray(demix_filter)
        if demix_filter.ndim == 2:
            demix_filter = np.tile(demix_filter, reps=(input.shape[1], 1, 1))
        return _lib.demix(input, demix_filter)

    def compute_negative_loglikelihood(self):
        raise NotImplementedError("Implement 'compute_negative_loglikelihood' function.")


class GaussIDLMA(IDLMAbase):
    """Drop-in for src/sss/idlma.py:88-258.

    normalize: 'projection-back' (the only normalisation `update_once` completes upstream; 'power', the default, raises there)
    threshold: bound on the condition number of W U below which an IP row update is accepted
    """

    def __init__(self, domain=2, normalize='power', reference_id=0, callback=None, dnn_flooring=1e-5, eps=EPS, threshold=THRESHOLD):
        super().__init__(normalize=normalize, callback=callback, dnn_flooring=dnn_flooring, eps=eps)
        assert 1 <= domain <= 2, "1 <= `domain` <= 2 is not satisfied."
        self.domain, self.reference_id, self.threshold = domain, reference_id, threshold

    def __call__(self, input, iteration=100, **kwargs):
        """input (n_channels, n_bins, n_frames) -> projection-backed estimates (n_sources, n_bins, n_frames)."""
        self._iterate(input, iteration, kwargs)
        output = self._estimates(projection_back=True)
        self._host['estimation'] = output
        return output

    def update_once(self, is_source_model_update=True):
        if is_source_model_update:
            self.update_source_model()
        self.update_space_model()
        # the reference checks `normalize` only now, after the sweep (src/sss/idlma.py:150-162)
        if not self.normalize:
            raise ValueError("Set normalize=True")
        if self.normalize != 'projection-back':
            raise ValueError("Not support normalization based on {}. Choose 'power' or 'projection-back'".format(self.normalize))
        self._handle.normalize()
        self._device_changed()

    def update_source_model(self):
        """Power of the current estimates -> DNN -> floored variances (src/sss/idlma.py:167-173)."""
        self._prepare()
        power = np.abs(self._estimates(projection_back=False))**2
        self.dnn_output = self.estimate_by_dnn(power)
        if self.dnn_flooring:
            self.floor_dnn_output()

    def update_space_model(self):
        """Weighted covariances and the gated IP sweep on the device (src/sss/idlma.py:175-210)."""
        self._prepare()
        self._send_variance()
        self._handle.update_once()
        self._device_changed()

    def _send_variance(self):
        """Upload R = dnn_output^(2/domain) (src/sss/idlma.py:181, :253) when it changed; the device applies the eps floor."""
        d = np.asarray(self.dnn_output)
        want = (self.n_sources, self.n_bins, self.n_frames)
        if d.shape != want:
            raise ValueError("dnn_output has shape {}, expected {}".format(d.shape, want))
        token = (id(self.dnn_output), float(d.sum()), float(self.domain))
        if self.__dict__.get('_variance_token') != token:
            R = d**(2 / self.domain)   # in the array's own precision, like the reference (float32 behind a torch DNN)
            self._handle.set_state(_lib.STATE_VARIANCE, np.ascontiguousarray(R, dtype=np.float64), np.float64)
            self.__dict__['_variance_token'] = token

    def estimate_by_dnn(self, input):
        """src/sss/idlma.py:212-226: the DNN sees power^(domain/2) and its output is raised to 2/domain.  `self.dnn` is a
        torch.nn.Module (float32 tensors, moved to the module's device) or any callable on NumPy arrays."""
        exponent = self.domain / 2
        if hasattr(self.dnn, 'parameters'):
            import torch
            with torch.no_grad():
                x = torch.Tensor(input**exponent)
                if next(self.dnn.parameters()).is_cuda:
                    x = x.cuda()
                y = self.dnn(x).cpu().numpy()
        else:
            y = np.asarray(self.dnn(input**exponent))
        return y**(2 / self.domain)

    def floor_dnn_output(self):
        self.dnn_output = np.maximum(self.dnn_output, self.dnn_flooring)

    def compute_negative_loglikelihood(self):
        """sum(P / R + log R) - 2 T sum_f log|det W_f| (src/sss/idlma.py:244-258), reduced on the device."""
        self._prepare()
        self._send_variance()
        return float(self._handle.loss()[0])
