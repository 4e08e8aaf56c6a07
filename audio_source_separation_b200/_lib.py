"""ctypes binding of libbssgpu.so (the C ABI declared in include/bssgpu.h).

There is no CPU path: if the shared library is missing, or the machine has no B200-class GPU, the
first use raises.  Nothing here imports `oracle/`.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('BSSGPU_LIBRARY') or os.path.join(_HERE, 'libbssgpu.so')   # BSSGPU_LIBRARY: A/B against another build

# enum bss_status
OK, EINVAL, ECUDA, ESINGULAR, ENOMEM, ESTATE, EUNSUPPORTED, ENCCL = 0, -1, -2, -3, -4, -5, -6, -7
# enum bss_method
GAUSS_ILRMA, T_ILRMA, AUX_LAPLACE_IVA, AUX_GAUSS_IVA, FAST_MNMF, IS_MNMF, GAUSS_IDLMA = 0, 1, 2, 3, 4, 5, 6
NMF_EUC, NMF_KL, NMF_IS, NMF_T, NMF_CAUCHY = 10, 11, 12, 13, 14
# enum bss_spatial / bss_normalize / bss_nmf_algorithm
SPATIAL_IP, SPATIAL_ISS, SPATIAL_IP2 = 0, 1, 2
NORMALIZE_NONE, NORMALIZE_POWER, NORMALIZE_PROJECTION_BACK = 0, 1, 2
ALG_MM, ALG_ME, ALG_NAIVE, ALG_MM_FAST = 0, 1, 2, 3
# enum bss_dtype
F32, F64, C64, C128, I32, I16 = 0, 1, 2, 3, 4, 5
# enum bss_state
(STATE_DEMIX_FILTER, STATE_ESTIMATION, STATE_BASIS, STATE_ACTIVATION, STATE_LATENT, STATE_DIAGONALIZER,
 STATE_SPATIAL, STATE_TARGET, STATE_COVARIANCE, STATE_GATE, STATE_VARIANCE, STATE_ORDER, STATE_EIGVAL) = range(13)
# enum bss_option / bss_info
OPT_IP_KERNEL, OPT_ACT_CHUNKS, OPT_BLOCKING_SYNC, OPT_SOURCE_MODEL, OPT_ASYNC_INPUT = 0, 1, 2, 3, 4
SOURCE_MODEL_AUTO, SOURCE_MODEL_THREE_PASS, SOURCE_MODEL_FUSED = 0, 1, 2
IP_AUTO, IP_THREAD_PER_BIN, IP_LANE_GROUP, IP_PAIRWISE = 0, 1, 2, 4
INFO_IP_KERNEL, INFO_GRAPH_REPLAYS, INFO_LAUNCHES, INFO_ACT_CHUNKS, INFO_SOURCE_MODEL = 0, 1, 2, 3, 4

_DTYPES = {np.dtype(np.float32): F32, np.dtype(np.float64): F64, np.dtype(np.complex64): C64,
           np.dtype(np.complex128): C128, np.dtype(np.int32): I32, np.dtype(np.int16): I16}


class Config(ctypes.Structure):
    _fields_ = [
        ('method', ctypes.c_int32), ('spatial', ctypes.c_int32), ('normalize', ctypes.c_int32),
        ('partitioning', ctypes.c_int32), ('algorithm', ctypes.c_int32), ('n_batch', ctypes.c_int32),
        ('n_channels', ctypes.c_int32), ('n_sources', ctypes.c_int32), ('n_bins', ctypes.c_int32),
        ('n_frames', ctypes.c_int32), ('n_basis', ctypes.c_int32), ('reference_id', ctypes.c_int32),
        ('device', ctypes.c_int32), ('stream_priority', ctypes.c_int32),
        ('domain', ctypes.c_double), ('nu', ctypes.c_double), ('eps', ctypes.c_double), ('threshold', ctypes.c_double),
    ]


# every symbol include/bssgpu.h declares: name -> (restype, argtypes)
_vp, _i, _d = ctypes.c_void_p, ctypes.c_int, ctypes.c_double
SIGNATURES = {
    'bss_create': (_i, [ctypes.POINTER(Config), ctypes.POINTER(_vp)]),
    'bss_destroy': (None, [_vp]),
    'bss_last_error': (ctypes.c_char_p, [_vp]),
    'bss_set_stream': (_i, [_vp, _vp]),
    'bss_synchronize': (_i, [_vp]),
    'bss_set_input': (_i, [_vp, _vp, _i]),
    'bss_set_input_waveform': (_i, [_vp, _vp, _i, _i, _i, _i, _vp]),
    'bss_stft_frames': (_i, [_i, _i, _i]),
    'bss_istft_length': (_i, [_i, _i, _i]),
    'bss_stft': (_i, [_i, _i, _i, _i, _i, _vp, _vp, _vp]),
    'bss_istft': (_i, [_i, _i, _i, _i, _i, _vp, _vp, _vp]),
    'bss_set_state': (_i, [_vp, _i, _vp, _i]),
    'bss_get_state': (_i, [_vp, _i, _vp, _i]),
    'bss_reset_spatial': (_i, [_vp]),
    'bss_set_update_pair': (_i, [_vp, _i, _i]),
    'bss_update_once': (_i, [_vp]),
    'bss_normalize': (_i, [_vp]),
    'bss_run': (_i, [_vp, _i]),
    'bss_run_record': (_i, [_vp, _i, ctypes.POINTER(_d)]),
    'bss_loss': (_i, [_vp, ctypes.POINTER(_d)]),
    'bss_separate': (_i, [_vp, _vp, _i, _i]),
    'bss_separate_device': (_i, [_vp, _vp, _i]),
    'bss_separate_waveform': (_i, [_vp, _vp, _i, _i, _i, _vp, _i]),
    'bss_separate_waveform_device': (_i, [_vp, _vp, _i, _i, _i, _vp, _i]),
    'bss_gather_outputs': (_i, [_vp, _vp, _i, _i, _vp, _vp, ctypes.c_size_t, ctypes.c_size_t]),
    'bss_peer_alloc': (_i, [_i, ctypes.c_size_t, ctypes.POINTER(_vp), _vp]),
    'bss_peer_open': (_i, [_i, _vp, ctypes.POINTER(_vp)]),
    'bss_peer_close': (_i, [_i, _vp]),
    'bss_peer_free': (_i, [_i, _vp]),
    'bss_push_outputs': (_i, [_vp, _i, _i, ctypes.POINTER(_vp), _vp, ctypes.c_size_t, ctypes.c_size_t]),
    'bss_compute_demix_filter': (_i, [_vp]),
    'bss_set_option': (_i, [_vp, _i, _i]),
    'bss_get_info': (_i, [_vp, _i, ctypes.POINTER(ctypes.c_int64)]),
    'bss_least_squares_map': (_i, [_i, _i, _i, _i, _i, _vp, _vp, _vp]),
    'bss_weighted_covariance': (_i, [_i, _i, _i, _i, _i, _vp, _vp, _vp]),
    'bss_ip_update': (_i, [_i, _i, _i, _vp, _vp, _vp, _d, _i, _d]),
    'bss_projection_back_scale': (_i, [_i, _i, _i, _i, _vp, _vp, _i, _vp]),
    'bss_demix': (_i, [_i, _i, _i, _i, _vp, _vp, _vp]),
    'bss_timer_begin': (_i, [_vp]),
    'bss_timer_end': (_i, [_vp, ctypes.POINTER(ctypes.c_float)]),
    'bss_time_covariance': (_i, [_vp, _i, ctypes.POINTER(ctypes.c_float)]),
    'bss_launch_count': (ctypes.c_int64, [_vp]),
    'bss_device_buffer': (_i, [_vp, _i, ctypes.POINTER(_vp), ctypes.POINTER(ctypes.c_size_t)]),
    'bss_version': (ctypes.c_char_p, []),
}

_lib = None


def load():
    """Load libbssgpu.so and bind every entry point (no CUDA call is made)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "libbssgpu.so is not built ({}). Run `python -c 'import __graft_entry__ as g; g.build()'` or "
            "`make -C audio_source_separation_b200/csrc`. There is no CPU implementation.".format(LIB_PATH))
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def _raise(code, message):
    if code == ESINGULAR:
        raise np.linalg.LinAlgError(message or "Singular matrix")
    if code == EINVAL:
        raise ValueError(message)
    if code == EUNSUPPORTED:
        raise NotImplementedError(message)
    if code == ENOMEM:
        raise MemoryError(message)
    if code == ENCCL:
        raise RuntimeError("NCCL: {}".format(message))
    raise RuntimeError("libbssgpu error {}: {}".format(code, message))


def check_static(code, what):
    if code != OK:
        _raise(code, "{} failed".format(what) + (": " + load().bss_last_error(None).decode() if code != ESINGULAR else ""))


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def as_host(a, dtype):
    """C-contiguous array of exactly `dtype` (no copy when it already is one)."""
    return np.ascontiguousarray(a, dtype=dtype)


class Handle:
    """Owns one bss_handle: a batch of B mixtures resident on one GPU."""

    def __init__(self, **cfg):
        lib = load()
        c = Config()
        defaults = dict(method=GAUSS_ILRMA, spatial=SPATIAL_IP, normalize=NORMALIZE_POWER, partitioning=0, algorithm=ALG_MM,
                        n_batch=1, n_channels=0, n_sources=0, n_bins=0, n_frames=0, n_basis=1, reference_id=0, device=0,
                        stream_priority=0, domain=2.0, nu=1.0, eps=1e-12, threshold=1e12)
        defaults.update(cfg)
        for k, v in defaults.items():
            setattr(c, k, v)
        self.cfg = defaults
        self._h = ctypes.c_void_p()
        self._lib = lib
        code = lib.bss_create(ctypes.byref(c), ctypes.byref(self._h))
        if code != OK:
            self._h = None
            _raise(code, lib.bss_last_error(None).decode())

    # shapes -------------------------------------------------------------------------------------
    @property
    def B(self):
        return self.cfg['n_batch']

    def _check(self, code):
        if code != OK:
            _raise(code, self._lib.bss_last_error(self._h).decode())

    def close(self):
        if getattr(self, '_h', None):
            self._lib.bss_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # data movement ------------------------------------------------------------------------------
    def set_stream(self, cuda_stream):
        self._check(self._lib.bss_set_stream(self._h, ctypes.c_void_p(cuda_stream)))

    def synchronize(self):
        self._check(self._lib.bss_synchronize(self._h))

    def set_input(self, x):
        if x.dtype != np.complex64:
            x = as_host(x, np.complex128)
        else:
            x = as_host(x, np.complex64)
        self._check(self._lib.bss_set_input(self._h, _ptr(x), _DTYPES[x.dtype]))

    def set_input_waveform(self, x, fft_size, hop_size, window):
        """x (B,C,n_samples) float32/float64 on the host; the STFT runs on the device."""
        x = as_host(x, x.dtype if x.dtype in (np.float32, np.int16) else np.float64)   # int16: PCM samples, scaled by 1 / 32768
        window = as_host(window, np.float64)
        self._check(self._lib.bss_set_input_waveform(self._h, _ptr(x), _DTYPES[x.dtype], x.shape[-1], int(fft_size), int(hop_size),
                                                     _ptr(window)))

    def set_input_waveform_ptr(self, ptr, dtype, n_samples, fft_size, hop_size, window):
        """Waveforms (B,C,n_samples) already sitting in (pinned) host memory at address `ptr`."""
        window = as_host(window, np.float64)
        self._check(self._lib.bss_set_input_waveform(self._h, ctypes.c_void_p(ptr), dtype, int(n_samples), int(fft_size), int(hop_size),
                                                     _ptr(window)))

    def set_input_ptr(self, ptr, dtype):
        """Input already sitting in (pinned) host memory at address `ptr`."""
        self._check(self._lib.bss_set_input(self._h, ctypes.c_void_p(ptr), dtype))

    def set_state(self, which, a, dtype):
        a = as_host(a, dtype)
        self._check(self._lib.bss_set_state(self._h, which, _ptr(a), _DTYPES[a.dtype]))

    def get_state(self, which, shape, dtype):
        out = np.empty(shape, dtype=dtype)
        self._check(self._lib.bss_get_state(self._h, which, _ptr(out), _DTYPES[out.dtype]))
        return out

    def reset_spatial(self):
        self._check(self._lib.bss_reset_spatial(self._h))

    # update loop --------------------------------------------------------------------------------
    def set_update_pair(self, m, n):
        self._check(self._lib.bss_set_update_pair(self._h, int(m), int(n)))

    def update_once(self):
        self._check(self._lib.bss_update_once(self._h))

    def normalize(self):
        self._check(self._lib.bss_normalize(self._h))

    def run(self, n_iter):
        self._check(self._lib.bss_run(self._h, int(n_iter)))

    def run_record(self, n_iter):
        """n_iter updates with the loss after each one; returns (n_iter, B)."""
        n_iter = int(n_iter)
        out = (ctypes.c_double * (max(n_iter, 1) * self.B))()
        self._check(self._lib.bss_run_record(self._h, n_iter, out))
        return np.array(out[:n_iter * self.B], dtype=np.float64).reshape(n_iter, self.B)

    def loss(self):
        out = (ctypes.c_double * self.B)()
        self._check(self._lib.bss_loss(self._h, out))
        return np.array(out[:], dtype=np.float64)

    def separate(self, shape, dtype=np.complex128, projection_back=True):
        out = np.empty(shape, dtype=dtype)
        self._check(self._lib.bss_separate(self._h, _ptr(out), _DTYPES[out.dtype], 1 if projection_back else 0))
        return out

    def separate_into(self, ptr, dtype, projection_back=True):
        self._check(self._lib.bss_separate(self._h, ctypes.c_void_p(ptr), dtype, 1 if projection_back else 0))

    def separate_waveform(self, n_signals_shape, fft_size, hop_size, window, dtype=np.float64, projection_back=True):
        """Separated estimates in the time domain, shape `n_signals_shape + (n_out,)` (ISTFT on the device)."""
        window = as_host(window, np.float64)
        n_out = self._lib.bss_istft_length(self.cfg['n_frames'], int(fft_size), int(hop_size))
        out = np.empty(tuple(n_signals_shape) + (n_out,), dtype=dtype)
        self._check(self._lib.bss_separate_waveform(self._h, _ptr(out), _DTYPES[out.dtype], int(fft_size), int(hop_size), _ptr(window),
                                                    1 if projection_back else 0))
        return out

    def separate_waveform_into(self, ptr, dtype, fft_size, hop_size, window, projection_back=True):
        window = as_host(window, np.float64)
        self._check(self._lib.bss_separate_waveform(self._h, ctypes.c_void_p(ptr), dtype, int(fft_size), int(hop_size), _ptr(window),
                                                    1 if projection_back else 0))

    def separate_waveform_device(self, device_ptr, dtype, fft_size, hop_size, window, projection_back=True):
        window = as_host(window, np.float64)
        self._check(self._lib.bss_separate_waveform_device(self._h, ctypes.c_void_p(device_ptr), dtype, int(fft_size), int(hop_size),
                                                           _ptr(window), 1 if projection_back else 0))

    def separate_device(self, device_ptr, projection_back=True):
        self._check(self._lib.bss_separate_device(self._h, ctypes.c_void_p(device_ptr), 1 if projection_back else 0))

    def compute_demix_filter(self):
        self._check(self._lib.bss_compute_demix_filter(self._h))

    def push_outputs(self, peers, send_ptr, dst_offset_bytes, nbytes):
        """Copy `nbytes` from send_ptr into every peer's result buffer (a `PeerBuffers`) at dst_offset_bytes, on the handle's stream."""
        self._check(self._lib.bss_push_outputs(self._h, peers.world, peers.rank, peers.bases, ctypes.c_void_p(send_ptr),
                                               int(dst_offset_bytes), int(nbytes)))

    def gather_outputs(self, comm, send_ptr, recv_base_ptr, nbytes, rank_stride_bytes):
        """All-gather on the handle's stream over `comm` (an `NcclComm`): rank r's `nbytes` land at recv_base + r * stride."""
        self._check(self._lib.bss_gather_outputs(self._h, comm.handle, comm.world, comm.rank, ctypes.c_void_p(send_ptr),
                                                 ctypes.c_void_p(recv_base_ptr), int(nbytes), int(rank_stride_bytes)))

    def set_option(self, option, value):
        self._check(self._lib.bss_set_option(self._h, int(option), int(value)))

    def get_info(self, what):
        out = ctypes.c_int64()
        self._check(self._lib.bss_get_info(self._h, int(what), ctypes.byref(out)))
        return int(out.value)

    # measurement --------------------------------------------------------------------------------
    def timer_begin(self):
        self._check(self._lib.bss_timer_begin(self._h))

    def timer_end(self):
        ms = ctypes.c_float()
        self._check(self._lib.bss_timer_end(self._h, ctypes.byref(ms)))
        return ms.value

    def time_covariance(self, repeat):
        ms = ctypes.c_float()
        self._check(self._lib.bss_time_covariance(self._h, int(repeat), ctypes.byref(ms)))
        return ms.value

    def launch_count(self):
        return int(self._lib.bss_launch_count(self._h))


class PeerBuffers:
    """One result buffer per rank of a node, each allocated by its owner (`bss_peer_alloc`) and mapped into every other
    process through CUDA IPC (`bss_peer_open`), so that a rank can push its outputs straight into its peers' buffers with
    device-to-device copies over NVLink.  The 64-byte handles travel through the caller's torch.distributed group."""

    def __init__(self, rank, world, device, nbytes, group=None):
        import torch.distributed as dist
        lib = load()
        self.rank, self.world, self.device, self.nbytes = int(rank), int(world), int(device), int(nbytes)
        self.own = ctypes.c_void_p()
        self._opened = []
        handle = (ctypes.c_ubyte * 64)()
        code = lib.bss_peer_alloc(self.device, self.nbytes, ctypes.byref(self.own), handle)
        if code != OK:
            self.own = ctypes.c_void_p()
        handles = [None] * self.world
        # every rank takes part in the exchange, also one whose allocation failed (it sends None and all ranks give up together)
        dist.all_gather_object(handles, bytes(handle) if code == OK else None, group=group)
        if any(hd is None for hd in handles):
            self.close()
            raise RuntimeError("bss_peer_alloc failed on rank(s) {}".format([r for r, hd in enumerate(handles) if hd is None]))
        self.bases = (ctypes.c_void_p * self.world)()
        self.bases[self.rank] = self.own.value
        for r in range(self.world):
            if r == self.rank:
                continue
            p = ctypes.c_void_p()
            buf = (ctypes.c_ubyte * 64).from_buffer_copy(handles[r])
            code = lib.bss_peer_open(self.device, buf, ctypes.byref(p))
            if code != OK:
                raise RuntimeError("bss_peer_open failed for rank {} ({})".format(r, code))
            self.bases[r] = p.value
            self._opened.append(p)

    def as_tensor(self, shape, dtype):
        """The rank's own buffer as a torch tensor (shares the memory; keep this object alive while the tensor is used)."""
        import torch
        typestr = {torch.float32: '<f4', torch.float64: '<f8'}[dtype]
        holder = type('DevicePointer', (), {})()
        holder.__cuda_array_interface__ = {'shape': tuple(shape), 'typestr': typestr, 'data': (int(self.own.value), False), 'version': 2}
        holder.owner = self
        return torch.as_tensor(holder, device=torch.device('cuda', self.device))

    def unmap(self):
        """Drop this process's mappings of the peers' buffers (before their owners free them)."""
        lib = load()
        for p in getattr(self, '_opened', []):
            lib.bss_peer_close(self.device, p)
        self._opened = []

    def close(self):
        self.unmap()
        if getattr(self, 'own', None) is not None and self.own.value:
            load().bss_peer_free(self.device, self.own)
            self.own = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class NcclComm:
    """An NCCL communicator of our own for `bss_gather_outputs` (one per process / GPU), created through the same
    libnccl.so.2 the process already carries.  The 128-byte unique id travels through the caller's torch.distributed group
    (any backend), which is only used for that."""

    def __init__(self, rank, world, device, group=None):
        import torch
        import torch.distributed as dist
        self.rank, self.world = int(rank), int(world)
        self._nccl = None
        for name in ('libnccl.so.2', 'libnccl.so'):
            try:
                self._nccl = ctypes.CDLL(name, mode=ctypes.RTLD_GLOBAL)
                break
            except OSError:
                continue
        if self._nccl is None:
            raise RuntimeError("libnccl.so.2 not found")
        class _Uid(ctypes.Structure):      # ncclUniqueId { char internal[128]; }, passed BY VALUE to ncclCommInitRank
            _fields_ = [('internal', ctypes.c_byte * 128)]
        uid = _Uid()
        if self.rank == 0:
            self._nccl.ncclGetUniqueId.argtypes = [ctypes.POINTER(_Uid)]
            self._nccl.ncclGetUniqueId.restype = ctypes.c_int
            rc = self._nccl.ncclGetUniqueId(ctypes.byref(uid))
            if rc != 0:
                raise RuntimeError("ncclGetUniqueId failed ({})".format(rc))
        box = [bytes(bytearray(uid.internal))]
        dist.broadcast_object_list(box, src=0, group=group)
        if len(box[0]) != 128:
            raise RuntimeError("NCCL unique id did not arrive")
        ctypes.memmove(ctypes.byref(uid), box[0], 128)
        torch.cuda.set_device(device)
        comm = ctypes.c_void_p()
        max_ctas = int(os.environ.get('BSSGPU_NCCL_MAX_CTAS', '0'))
        rc = -1
        if max_ctas > 0:
            # ncclConfig_t as of NCCL 2.27 (newer libraries accept an older, shorter struct): cap the CTAs the collective's
            # kernel may occupy, so that a gather running beside the update loops takes few SMs from them
            class _Config(ctypes.Structure):
                _fields_ = [('size', ctypes.c_size_t), ('magic', ctypes.c_uint), ('version', ctypes.c_uint), ('blocking', ctypes.c_int),
                            ('cgaClusterSize', ctypes.c_int), ('minCTAs', ctypes.c_int), ('maxCTAs', ctypes.c_int),
                            ('netName', ctypes.c_char_p), ('splitShare', ctypes.c_int), ('trafficClass', ctypes.c_int),
                            ('commName', ctypes.c_char_p), ('collnetEnable', ctypes.c_int), ('CTAPolicy', ctypes.c_int),
                            ('shrinkShare', ctypes.c_int), ('nvlsCTAs', ctypes.c_int)]
            undef = -2 ** 31
            cfg = _Config(ctypes.sizeof(_Config), 0xcafebeef, 22703, undef, undef, undef, max_ctas, None, undef, undef, None, undef, undef,
                          undef, undef)
            try:
                fn = self._nccl.ncclCommInitRankConfig
                fn.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_int, _Uid, ctypes.c_int, ctypes.POINTER(_Config)]
                fn.restype = ctypes.c_int
                rc = fn(ctypes.byref(comm), self.world, uid, self.rank, ctypes.byref(cfg))
            except AttributeError:
                rc = -1
        if rc != 0:
            self._nccl.ncclCommInitRank.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_int, _Uid, ctypes.c_int]
            self._nccl.ncclCommInitRank.restype = ctypes.c_int
            rc = self._nccl.ncclCommInitRank(ctypes.byref(comm), self.world, uid, self.rank)
        if rc != 0:
            raise RuntimeError("ncclCommInitRank failed ({})".format(rc))
        self.handle = comm

    def close(self):
        if getattr(self, 'handle', None):
            self._nccl.ncclCommDestroy.argtypes = [ctypes.c_void_p]
            self._nccl.ncclCommDestroy(self.handle)
            self.handle = None


# stateless primitives ---------------------------------------------------------------------------

def weighted_covariance(x, r, device=0):
    """U[n,f] = mean_t x x^H / r[n,f,t]; x (C,F,T), r broadcastable to (N,F,T) -> (N,F,C,C) complex128."""
    lib = load()
    x = as_host(x, np.complex128)
    C, F, T = x.shape
    r = np.ascontiguousarray(np.broadcast_to(np.asarray(r, dtype=np.float64), (C, F, T)))
    u = np.empty((C, F, C, C), dtype=np.complex128)
    code = lib.bss_weighted_covariance(device, C, C, F, T, _ptr(x), _ptr(r), _ptr(u))
    check_static(code, 'bss_weighted_covariance')
    return u


def ip_update(w, u, threshold=1e12, floor_den=False, eps=1e-12, device=0):
    """Gauss-Seidel IP sweep; returns (W_new (F,N,C), gate (N,F) bool)."""
    lib = load()
    w = np.array(w, dtype=np.complex128, order='C', copy=True)
    u = as_host(u, np.complex128)
    F, N, C = w.shape
    gate = np.empty((N, F), dtype=np.int32)
    code = lib.bss_ip_update(device, C, F, _ptr(w), _ptr(u), _ptr(gate), float(threshold), 1 if floor_den else 0, float(eps))
    check_static(code, 'bss_ip_update')
    return w, gate.astype(bool)


def projection_back_scale(x, w, reference_id=0, device=0):
    lib = load()
    x = as_host(x, np.complex128)
    w = as_host(w, np.complex128)
    C, F, T = x.shape
    scale = np.empty((C, F), dtype=np.complex128)
    code = lib.bss_projection_back_scale(device, C, F, T, _ptr(x), _ptr(w), int(reference_id), _ptr(scale))
    check_static(code, 'bss_projection_back_scale')
    return scale


def least_squares_map(a, b, device=0):
    """Per-bin M_f = A_f B_f^H (B_f B_f^H)^-1; a (Ra,F,T), b (Rb,F,T) -> (Ra,Rb,F) complex128."""
    lib = load()
    a = as_host(a, np.complex128)
    b = as_host(b, np.complex128)
    if a.ndim != 3 or b.ndim != 3 or a.shape[1:] != b.shape[1:]:
        raise ValueError("expected (rows, n_bins, n_frames) arrays with equal bins and frames, got {} and {}".format(a.shape, b.shape))
    Ra, F, T = a.shape
    Rb = b.shape[0]
    out = np.empty((Ra, Rb, F), dtype=np.complex128)
    code = lib.bss_least_squares_map(device, Ra, Rb, F, T, _ptr(a), _ptr(b), _ptr(out))
    check_static(code, 'bss_least_squares_map')
    return out


def demix(x, w, device=0):
    """Y[n,f,t] = sum_c W[f,n,c] X[c,f,t] on the GPU (complex64 arithmetic), returned as complex128."""
    lib = load()
    x = as_host(x, np.complex128)
    w = as_host(w, np.complex128)
    C, F, T = x.shape
    y = np.empty((w.shape[1], F, T), dtype=np.complex128)
    code = lib.bss_demix(device, C, F, T, 0, _ptr(x), _ptr(w), _ptr(y))
    check_static(code, 'bss_demix')
    return y


def stft_frames(n_samples, fft_size, hop_size):
    n = load().bss_stft_frames(int(n_samples), int(fft_size), int(hop_size))
    if n < 0:
        _raise(n, "invalid STFT geometry")
    return n


def istft_length(n_frames, fft_size, hop_size):
    n = load().bss_istft_length(int(n_frames), int(fft_size), int(hop_size))
    if n < 1:
        _raise(EINVAL, "invalid ISTFT geometry")
    return n


def stft(x, fft_size, hop_size, window, device=0):
    """x (..., n_samples) real -> (..., fft_size // 2 + 1, n_frames) complex128 (scipy.signal.stft semantics)."""
    lib = load()
    x = np.asarray(x)
    lead = x.shape[:-1]
    x2 = as_host(x.reshape(-1, x.shape[-1]), np.float64)
    window = as_host(window, np.float64)
    n_frames = stft_frames(x2.shape[1], fft_size, hop_size)
    out = np.empty((x2.shape[0], fft_size // 2 + 1, n_frames), dtype=np.complex128)
    code = lib.bss_stft(device, x2.shape[0], x2.shape[1], int(fft_size), int(hop_size), _ptr(window), _ptr(x2), _ptr(out))
    check_static(code, 'bss_stft')
    return out.reshape(lead + out.shape[1:])


def istft(z, fft_size, hop_size, window, device=0):
    """z (..., fft_size // 2 + 1, n_frames) complex -> (..., n_out) float64 (scipy.signal.istft semantics)."""
    lib = load()
    z = np.asarray(z)
    lead = z.shape[:-2]
    z2 = as_host(z.reshape((-1,) + z.shape[-2:]), np.complex128)
    if z2.shape[1] != fft_size // 2 + 1:
        raise ValueError("expected {} bins, got {}".format(fft_size // 2 + 1, z2.shape[1]))
    window = as_host(window, np.float64)
    n_out = lib.bss_istft_length(z2.shape[2], int(fft_size), int(hop_size))
    if n_out < 1:
        _raise(EINVAL, "invalid ISTFT geometry")
    out = np.empty((z2.shape[0], n_out), dtype=np.float64)
    code = lib.bss_istft(device, z2.shape[0], z2.shape[2], int(fft_size), int(hop_size), _ptr(window), _ptr(z2), _ptr(out))
    check_static(code, 'bss_istft')
    return out.reshape(lead + (n_out,))
