"""Single-channel NMF on the GPU with the reference's class surface (src/algorithm/nmf.py).

`EUCNMF` (:150), `KLNMF` (:209), `ISNMF` (:268), `tNMF` (:358) and `CauchyNMF` (:430): same constructor
arguments, `__call__(target, iteration)`, `update`, `update_once`, `loss`, `basis`, `activation`.  The
multiplicative updates and the per-iteration criterion run in CUDA kernels (fp64) behind libbssgpu's C ABI;
`update(iteration)` queues the whole loop on the device and fetches the loss history once.

As in the reference, `_reset` ALWAYS draws a fresh `basis` (n_bins, n_basis) and then `activation`
(n_basis, n_frames) from NumPy's global legacy RNG (src/algorithm/nmf.py:42-43): presets passed through
kwargs are overwritten.  Assign `model.basis` / `model.activation` after `_reset()` to inject a state.
"""
import numpy as np

from .. import _lib
from .._model import DeviceModel

EPS = 1e-12

__metrics__ = ['EUC', 'KL', 'IS']

_ALGORITHMS = {'mm': _lib.ALG_MM, 'me': _lib.ALG_ME, 'naive-multipricative': _lib.ALG_NAIVE, 'mm_fast': _lib.ALG_MM_FAST}


class NMFbase(DeviceModel):
    """src/algorithm/nmf.py:10-56"""

    _STATE_IDS = {'basis': _lib.STATE_BASIS, 'activation': _lib.STATE_ACTIVATION}
    _method = None
    _supported = ('mm',)

    def __init__(self, n_basis=2, eps=EPS):
        """
        Args:
            n_basis: number of basis
        """
        DeviceModel.__init__(self)
        self.n_basis = n_basis
        self.loss = []

        self.eps = eps
        self.target = None

    def __call__(self, target, iteration=100, **kwargs):
        self.target = target

        self._reset(**kwargs)

        self.update(iteration=iteration)

        T, V = self.basis, self.activation

        return T.copy(), V.copy()

    # -- device plumbing -----------------------------------------------------------------------------
    def _state_shape(self, name):
        n_bins, n_frames = self.target.shape
        if name == 'basis':
            return (n_bins, self.n_basis)
        if name == 'activation':
            return (self.n_basis, n_frames)
        raise KeyError(name)

    def _config(self):
        algorithm = getattr(self, 'algorithm', 'mm')
        if algorithm not in self._supported:
            raise ValueError("Not support {} based update.".format(algorithm))
        n_bins, n_frames = self.target.shape
        return dict(method=self._method, algorithm=_ALGORITHMS[algorithm], n_batch=1, n_channels=1, n_sources=1, n_bins=n_bins,
                    n_frames=n_frames, n_basis=self.n_basis, domain=float(getattr(self, 'domain', 2)),
                    nu=float(getattr(self, 'nu', 1)), eps=float(self.eps))

    def _prepare(self):
        assert self.target is not None, "Specify data!"
        self._check_arguments()
        cfg = self._config()
        self._open_handle(tuple(sorted(cfg.items())), **cfg)
        target = self.target
        token = (id(target), target.shape, target.dtype.str, self._fingerprint(np.asarray(target)))
        if self._input_token != token:
            self._handle.set_state(_lib.STATE_TARGET, target, np.float64)
            self.__dict__['_input_token'] = token
        self._push()
        self._on_device.update(('basis', 'activation'))

    def _check_arguments(self):
        pass

    # -- reference surface -----------------------------------------------------------------------------
    def _reset(self, **kwargs):
        assert self.target is not None, "Specify data!"

        for key in kwargs.keys():
            setattr(self, key, kwargs[key])

        n_basis = self.n_basis
        n_bins, n_frames = self.target.shape

        self.basis = np.random.rand(n_bins, n_basis)
        self.activation = np.random.rand(n_basis, n_frames)
        self.__dict__['_input_token'] = None   # every __call__ re-reads the target, like the reference

    def update(self, iteration=100):
        """`iteration` x (update_once, criterion) without leaving the device (src/algorithm/nmf.py:165-174)."""
        self._prepare()
        loss = self._handle.run_record(iteration)
        self._device_changed()
        self.loss.extend(float(v) for v in loss[:, 0])

    def update_once(self):
        self._prepare()
        self._handle.update_once()
        self._device_changed()


class EUCNMF(NMFbase):
    """src/algorithm/nmf.py:150-207"""
    _method = _lib.NMF_EUC

    def __init__(self, n_basis=2, domain=2, algorithm='mm', eps=EPS):
        """
        Args:
            n_basis: number of basis
        """
        super().__init__(n_basis=n_basis, eps=eps)

        assert 1 <= domain <= 2, "1 <= `domain` <= 2 is not satisfied."
        assert algorithm == 'mm', "algorithm must be 'mm'."

        self.domain = domain
        self.algorithm = algorithm
        self.criterion = lambda input, target: (target - input)**2


class KLNMF(NMFbase):
    """src/algorithm/nmf.py:209-266"""
    _method = _lib.NMF_KL

    def __init__(self, n_basis=2, domain=2, algorithm='mm', eps=EPS):
        """
        Args:
            K: number of basis
        """
        super().__init__(n_basis=n_basis, eps=eps)

        assert 1 <= domain <= 2, "1 <= `domain` <= 2 is not satisfied."
        assert algorithm == 'mm', "algorithm must be 'mm'."

        self.domain = domain
        self.algorithm = algorithm

    @staticmethod
    def criterion(input, target, eps=EPS):
        """src/criterion/divergence.py:34-45"""
        _input, _target = input + eps, target + eps
        return _target * np.log(_target / _input) + _input - _target


class ISNMF(NMFbase):
    """src/algorithm/nmf.py:268-356"""
    _method = _lib.NMF_IS
    _supported = ('mm', 'me')

    def __init__(self, n_basis=2, domain=2, algorithm='mm', eps=EPS):
        """
        Args:
            K: number of basis
            algorithm: 'mm': MM algorithm based update, 'me': ME algorithm based update
        """
        super().__init__(n_basis=n_basis, eps=eps)

        assert 1 <= domain <= 2, "1 <= `domain` <= 2 is not satisfied."

        self.domain = domain
        self.algorithm = algorithm

    def _check_arguments(self):
        if self.algorithm == 'me':
            assert self.domain == 2, "Only domain = 2 is supported."

    @staticmethod
    def criterion(input, target, eps=EPS):
        """src/criterion/divergence.py:21-32"""
        ratio = (target + eps) / (input + eps)
        return ratio - np.log(ratio) - 1


class tNMF(NMFbase):
    """src/algorithm/nmf.py:358-428"""
    _method = _lib.NMF_T

    def __init__(self, n_basis=2, nu=1e+3, domain=2, algorithm='mm', eps=EPS):
        """
        Args:
            K: number of basis
            algorithm: 'mm': MM algorithm based update
        """
        super().__init__(n_basis=n_basis, eps=eps)

        def t_divergence(input, target):
            _input, _target = input + eps, target + eps

            return np.log(_input) + (2 + self.nu) / 2 * np.log(1 + (2 / nu) * (_target / _input))

        assert 1 <= domain <= 2, "1 <= `domain` <= 2 is not satisfied."

        self.nu = nu
        self.domain = domain
        self.algorithm = algorithm
        self.criterion = t_divergence

    def _check_arguments(self):
        if self.algorithm == 'mm':
            assert self.domain == 2, "`domain` is expected 2."


class CauchyNMF(NMFbase):
    """src/algorithm/nmf.py:430-595"""
    _method = _lib.NMF_CAUCHY
    _supported = ('naive-multipricative', 'mm', 'me', 'mm_fast')

    def __init__(self, n_basis, domain=2, algorithm='naive-multipricative', eps=EPS):
        super().__init__(n_basis=n_basis, eps=eps)

        def cauchy_divergence(input, target):
            eps = self.eps

            _input, _target = input + eps, target + eps
            numerator = 2 * _target**2 + _input**2
            denominator = 3 * _target**2

            return np.log(_target / _input) + (3 / 2) * np.log(numerator / denominator)

        assert domain == 2, "Only `domain` = 2 is supported."

        self.domain = domain
        self.algorithm = algorithm
        self.criterion = cauchy_divergence

    def _check_arguments(self):
        assert self.domain == 2, "Only 'domain' = 2 is supported."
