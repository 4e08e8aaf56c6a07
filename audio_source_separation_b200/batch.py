"""Batches of independent mixtures (our extension: the reference has no batch axis, one model == one mixture).

`BatchedGaussILRMA` runs B Gauss-ILRMA problems of identical shape side by side on one GPU: every kernel
takes the mixture index as an extra grid dimension, nothing is shared between mixtures, and the results
are identical to B separate `GaussILRMA` runs.  `shard_range` / `gather_outputs` split a batch over the
ranks of a torch.distributed job (one process per GPU); the only communication is the final all-gather of
the separated outputs over NCCL.
"""
import os

import numpy as np

from . import _lib
from ._model import parse_spatial, parse_normalize

EPS = 1e-12
THRESHOLD = 1e+12


def _check_presets(B, C, F, T, K, demix_filter=None, basis=None, activation=None):
    """The C ABI carries no lengths: a wrongly shaped preset would make the library read past the end of the host buffer."""
    for name, value, shape in (('demix_filter', demix_filter, (B, F, C, C)), ('basis', basis, (B, C, F, K)),
                               ('activation', activation, (B, C, K, T))):
        if value is not None and tuple(np.shape(value)) != shape:
            raise ValueError("{} has shape {}, expected {}".format(name, tuple(np.shape(value)), shape))


def _flag_device(device, group):
    """Where the one-element agreement tensors of the sharded job live: on the GPU, in host memory for gloo groups."""
    import torch
    import torch.distributed as dist
    return torch.device('cpu') if str(dist.get_backend(group)) == 'gloo' else torch.device('cuda', device)


class BatchedGaussILRMA:
    def __init__(self, n_basis=10, domain=2, normalize='power', algorithm_spatial='IP', reference_id=0, eps=EPS,
                 threshold=THRESHOLD, device=0):
        assert 1 <= domain <= 2, "1 <= `domain` <= 2 is not satisfied."
        self.n_basis = n_basis
        self.domain = domain
        self.normalize = normalize
        self.algorithm_spatial = algorithm_spatial
        self.reference_id = reference_id
        self.eps = eps
        self.threshold = threshold
        self.device = device
        self.handle = None
        self._key = None

    def open(self, B, C, F, T):
        key = (B, C, F, T)
        if self.handle is not None and self._key == key:
            return self.handle
        if self.handle is not None:
            self.handle.close()
        self.handle = _lib.Handle(method=_lib.GAUSS_ILRMA, spatial=parse_spatial(self.algorithm_spatial),
                                  normalize=parse_normalize(self.normalize), n_batch=B, n_channels=C, n_sources=C, n_bins=F,
                                  n_frames=T, n_basis=self.n_basis, reference_id=self.reference_id, device=self.device,
                                  domain=float(self.domain), eps=float(self.eps), threshold=float(self.threshold))
        self._key = key
        self.shape = key
        return self.handle

    def reset(self, X, demix_filter=None, basis=None, activation=None):
        """Upload a batch (B,C,F,T) and its initial state; missing state is drawn like the reference's `_reset`
        (np.random.rand: basis (B,N,F,K) first, then activation (B,N,K,T))."""
        B, C, F, T = X.shape
        K = self.n_basis
        _check_presets(B, C, F, T, K, demix_filter, basis, activation)
        h = self.open(B, C, F, T)
        h.set_input(X)
        if demix_filter is None:
            h.reset_spatial()
        else:
            h.set_state(_lib.STATE_DEMIX_FILTER, demix_filter, np.complex128)
        if basis is None:
            basis = np.random.rand(B, C, F, K)
        if activation is None:
            activation = np.random.rand(B, C, K, T)
        h.set_state(_lib.STATE_BASIS, basis, np.float64)
        h.set_state(_lib.STATE_ACTIVATION, activation, np.float64)
        return h

    def __call__(self, X, iteration=100, dtype=np.complex128, **presets):
        """X (B,C,F,T) -> projection-backed estimates (B,N,F,T)."""
        h = self.reset(X, **presets)
        h.run(iteration)
        B, C, F, T = X.shape
        return h.separate((B, C, F, T), dtype, projection_back=True)

    def _pipelined(self, B, C, F, T, iteration, basis, activation, pipeline, feed, drain, on_done=None):
        """Run `feed(h, lo, hi)` -> `iteration` updates -> `drain(h, lo, hi)` for every sub-batch [lo, hi) of a batch of B
        mixtures, each sub-batch on its own handle, CUDA stream and host thread, so that the copies of one sub-batch overlap
        the update loop of the others.  `pipeline`: a count, a list of sub-batch sizes, or 'ramp' (`ramp_sizes(B)`).
        `on_done(i, lo, hi)` is called on the calling thread for every sub-batch, in index order, as soon as it has drained
        (`self._parts[i][1]` is its handle)."""
        from concurrent.futures import ThreadPoolExecutor
        K = self.n_basis
        if basis is None:
            basis = np.random.rand(B, C, F, K)
        if activation is None:
            activation = np.random.rand(B, C, K, T)
        if pipeline == 'ramp':
            pipeline = ramp_sizes(B)
        if isinstance(pipeline, (list, tuple)):
            # explicit sub-batch sizes, e.g. small first and last ones: the first upload and the last download are the only
            # copies nothing overlaps
            sizes = [int(v) for v in pipeline if int(v) > 0]
            if sum(sizes) != B:
                raise ValueError("pipeline sizes {} do not add up to the batch size {}".format(sizes, B))
            n_parts = len(sizes)
            edges = np.concatenate(([0], np.cumsum(sizes)))
            spans = [(int(edges[i]), int(edges[i + 1])) for i in range(n_parts)]
        else:
            n_parts = max(1, min(int(pipeline), B))
            spans = [shard_range(B, i, n_parts) for i in range(n_parts)]
        if not hasattr(self, '_parts') or len(self._parts) != n_parts:
            for slot in getattr(self, '_parts', []):
                if slot is not None:
                    slot[1].close()
            self._parts = [None] * n_parts

        import threading
        import time
        t_start = time.perf_counter()
        # Inputs go up in index order: the host link carries one copy at a time, and the sub-batch whose input arrives first
        # is the one whose update loop can start (and whose outputs can leave) first.  Left to race, the host threads put the
        # large middle sub-batches last at 8 GPUs per node -- their small state uploads queue behind the other threads' input
        # copies -- and most of the batch then finishes together at the very end (profiles/round2_scaling.md).
        fed = [threading.Event() for _ in spans]
        # (Queueing ALL inputs before the first update loop is launched was measured too: every sub-batch then finishes at the
        # end together, 50.7 against 44.2 ms at 4 GPUs.)
        marks = [dict(size=hi - lo) for lo, hi in spans]
        self.timeline = marks   # per sub-batch: ms since the start of the call at which each phase returned to the host

        def job(i):
            lo, hi = spans[i]
            key = (hi - lo, C, F, T)
            slot = self._parts[i]
            if slot is None or slot[0] != key:
                if slot is not None:
                    slot[1].close()
                h = _lib.Handle(method=_lib.GAUSS_ILRMA, spatial=parse_spatial(self.algorithm_spatial),
                                normalize=parse_normalize(self.normalize), n_batch=hi - lo, n_channels=C, n_sources=C, n_bins=F,
                                n_frames=T, n_basis=K, reference_id=self.reference_id, device=self.device,
                                domain=float(self.domain), eps=float(self.eps), threshold=float(self.threshold),
                                stream_priority=-(n_parts - 1 - i))   # earlier sub-batches finish first: their D2H overlaps the rest
                # several host threads per GPU (and several processes per node) wait on their streams at the same time:
                # sleep instead of spinning, the threads that still have launches to issue need the cores
                h.set_option(_lib.OPT_BLOCKING_SYNC, 1 if n_parts > 1 else 0)
                # the input copy is only queued by `feed` (the caller's arrays live until this call returns, and it returns
                # after every sub-batch has drained): the copies of all sub-batches go over the host link back to back
                h.set_option(_lib.OPT_ASYNC_INPUT, 1)
                self._parts[i] = slot = (key, h)
            h = slot[1]
            try:
                h.reset_spatial()
                h.set_state(_lib.STATE_BASIS, basis[lo:hi], np.float64)
                h.set_state(_lib.STATE_ACTIVATION, activation[lo:hi], np.float64)
                marks[i]['state'] = round(1e3 * (time.perf_counter() - t_start), 2)
                if i > 0:
                    fed[i - 1].wait()
                feed(h, lo, hi)     # queues the input copy (+ STFT) on the sub-batch's stream
            finally:
                fed[i].set()
            marks[i]['input'] = round(1e3 * (time.perf_counter() - t_start), 2)
            h.run(iteration)
            marks[i]['queued'] = round(1e3 * (time.perf_counter() - t_start), 2)
            drain(h, lo, hi)
            marks[i]['out'] = round(1e3 * (time.perf_counter() - t_start), 2)
            return h.launch_count()

        if n_parts == 1:
            job(0)
            if on_done is not None:
                on_done(0, *spans[0])
        else:
            def guarded(i):
                try:
                    return job(i)
                finally:
                    fed[i].set()    # also when the sub-batch failed before its upload: the next one must not wait for ever

            with ThreadPoolExecutor(max_workers=n_parts) as pool:
                futures = [pool.submit(guarded, i) for i in range(n_parts)]
                for i, fut in enumerate(futures):
                    fut.result()
                    if on_done is not None:
                        on_done(i, *spans[i])

    def separate_batch(self, X, out=None, iteration=100, basis=None, activation=None, pipeline='ramp', device_out=None):
        """Whole job for a batch held in host memory: X (B,C,F,T) complex64/128 -> projection-backed estimates written to
        `out` (B,N,F,T) complex64 (allocated when None).  The batch is cut into `pipeline` sub-batches (a count, a list of
        sub-batch sizes, or 'ramp' = `ramp_sizes(B)`), each with its own
        handle and CUDA stream and driven by its own host thread, so the host->device copy of one sub-batch and the
        device->host copy of another overlap the update loop of the rest (pass pinned arrays to make the copies
        asynchronous).  Mixtures are independent, so the result is identical to one undivided call.
        `device_out`: address of a device buffer (B,N,F,T) complex64 on this model's GPU; the estimates are left there
        instead of being copied to the host (`separate_batch_sharded` gathers them over NCCL), and None is returned."""
        B, C, F, T = X.shape
        _check_presets(B, C, F, T, self.n_basis, None, basis, activation)
        if device_out is None:
            if out is None:
                out = np.empty((B, C, F, T), dtype=np.complex64)
            if not (out.shape == (B, C, F, T) and out.dtype == np.complex64 and out.flags.c_contiguous):
                raise ValueError("out must be a C-contiguous complex64 array of shape {}".format((B, C, F, T)))
        X = X if X.flags.c_contiguous and X.dtype in (np.complex64, np.complex128) else np.ascontiguousarray(X, np.complex128)
        x_dtype = _lib.C64 if X.dtype == np.complex64 else _lib.C128

        def feed(h, lo, hi):
            h.set_input_ptr(X[lo:hi].ctypes.data, x_dtype)

        def drain(h, lo, hi):
            if device_out is not None:
                h.separate_device(int(device_out) + lo * C * F * T * 8, projection_back=True)
                h.synchronize()
            else:
                h.separate_into(out[lo:hi].ctypes.data, _lib.C64, projection_back=True)

        self._pipelined(B, C, F, T, iteration, basis, activation, pipeline, feed, drain)
        return out if device_out is None else None

    def separate_waveform_batch(self, x, fft_size, hop_size=None, window_fn='hann', out=None, iteration=100, basis=None,
                                activation=None, pipeline='ramp', device_out=None, loss_out=None, on_done=None):
        """The whole job in the time domain, pipelined like `separate_batch`: x (B,C,n_samples) float32/float64 in host
        memory -> separated signals (B,N,n_out) of the same dtype written to `out` (allocated when None), n_out = the length
        scipy.signal.istft returns.  x may also be int16 PCM (a quarter of the bytes of the spectrograms): the samples are
        scaled by 1 / 32768 on the device, as the reference's notebooks do after reading a wav file, and the output is float32.  STFT (src/transform/stft.py:4-8), update loop, projection back and ISTFT (:10-17) run on
        the device, so only waveforms cross PCIe: half the bytes of the spectrograms at 50 % overlap.
        `device_out`: address of a device buffer (B,N,n_out) of x's dtype on this model's GPU: the separated signals are left
        there instead of being copied to the host (None is returned).  `loss_out`: a float64 array (B,) that receives the
        final negative log-likelihood of every mixture (one small device-to-host read per sub-batch)."""
        from scipy import signal as ss
        if x.dtype not in (np.float32, np.float64, np.int16) or not x.flags.c_contiguous:
            x = np.ascontiguousarray(x, dtype=np.float64)
        out_dtype = np.dtype(np.float32) if x.dtype == np.int16 else x.dtype
        B, C, n_samples = x.shape
        if hop_size is None:
            hop_size = fft_size // 2
        window = np.ascontiguousarray(ss.get_window(window_fn, fft_size), dtype=np.float64)
        F, T = fft_size // 2 + 1, _lib.stft_frames(n_samples, fft_size, hop_size)
        n_out = _lib.istft_length(T, fft_size, hop_size)
        _check_presets(B, C, F, T, self.n_basis, None, basis, activation)
        if device_out is None:
            if out is None:
                out = np.empty((B, C, n_out), dtype=out_dtype)
            if not (out.shape == (B, C, n_out) and out.dtype == out_dtype and out.flags.c_contiguous):
                raise ValueError("out must be a C-contiguous {} array of shape {}".format(out_dtype, (B, C, n_out)))
        if loss_out is not None and not (loss_out.shape == (B,) and loss_out.dtype == np.float64):
            raise ValueError("loss_out must be a float64 array of shape {}".format((B,)))
        in_dtype = {np.dtype(np.float32): _lib.F32, np.dtype(np.float64): _lib.F64, np.dtype(np.int16): _lib.I16}[x.dtype]
        dtype = _lib.F32 if out_dtype == np.float32 else _lib.F64
        esz = out_dtype.itemsize

        def feed(h, lo, hi):
            h.set_input_waveform_ptr(x[lo:hi].ctypes.data, in_dtype, n_samples, fft_size, hop_size, window)

        def drain(h, lo, hi):
            if device_out is not None:
                h.separate_waveform_device(int(device_out) + lo * C * n_out * esz, dtype, fft_size, hop_size, window, projection_back=True)
                if loss_out is None:
                    h.synchronize()
            else:
                h.separate_waveform_into(out[lo:hi].ctypes.data, dtype, fft_size, hop_size, window, projection_back=True)
            if loss_out is not None:
                loss_out[lo:hi] = h.loss()   # waits for the stream: everything queued above is complete afterwards

        self._pipelined(B, C, F, T, iteration, basis, activation, pipeline, feed, drain, on_done=on_done)
        return out if device_out is None else None

    def separate_waveform_batch_sharded(self, x, fft_size, hop_size=None, window_fn='hann', iteration=100, basis=None, activation=None,
                                        group=None, pipeline='ramp', loss_out=None, overlap_gather=True, copy=False):
        """BASELINE configs[4] as one call, time domain in / time domain out (one process per GPU, torch.distributed
        initialised by the caller; without it: one GPU, no collective).  Every rank passes the SAME global description --
        x (B,C,n_samples) float32/float64 in host memory, of which it reads only its own contiguous shard
        `shard_range(B, rank, world)` -- uploads its shard pipelined against the update loop (STFT, `iteration` updates,
        projection back and ISTFT on the device), leaves its separated signals on its GPU and takes part in the one
        exchange of the path: the all-gather of the separated outputs over NVLink -- by default pushed into the peers' result
        buffers through peer memory (`bss_push_outputs`), with BSSGPU_GATHER_MODE=nccl the NCCL collective
        (`bss_gather_outputs`).  Returns a torch tensor (B,N,n_out) of x's dtype on this rank's GPU with the signals of ALL
        mixtures in batch order; in the peer-memory form it is a view of a buffer the model keeps and the next job on ANY
        rank overwrites (`copy=True` returns a tensor of its own instead, one device-to-device copy).  `loss_out` (B_local,) float64 receives the final losses of this rank's mixtures (the job's
        device-to-host read)."""
        import torch
        import torch.distributed as dist
        world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        rank = dist.get_rank(group) if world > 1 else 0
        B, C, n_samples = x.shape
        if hop_size is None:
            hop_size = fft_size // 2
        lo, hi = shard_range(B, rank, world)
        T = _lib.stft_frames(n_samples, fft_size, hop_size)
        n_out = _lib.istft_length(T, fft_size, hop_size)
        xl = x[lo:hi]
        tdtype = torch.float64 if xl.dtype == np.float64 else torch.float32
        device = torch.device('cuda', self.device)
        sizes = [shard_range(B, r, world)[1] - shard_range(B, r, world)[0] for r in range(world)]
        even = world > 1 and len(set(sizes)) == 1
        if world == 1 or not even:
            y_local = torch.empty((hi - lo, C, n_out), dtype=tdtype, device=device)
            if hi > lo:
                self.separate_waveform_batch(xl, fft_size, hop_size, window_fn, iteration=iteration,
                                             basis=None if basis is None else basis[lo:hi],
                                             activation=None if activation is None else activation[lo:hi], pipeline=pipeline,
                                             device_out=y_local.data_ptr(), loss_out=loss_out)
            if world == 1:
                return y_local
            return gather_outputs(y_local, world, group=group, sizes=sizes)
        # equal shards: every rank cuts its shard into the same sub-batches, and the outputs of sub-batch i are gathered (into
        # their final places, batch order) as soon as every rank has finished it -- the collective overlaps the update loops
        # of the later sub-batches; only the gather of the last, small sub-batch is exposed.  NCCL calls are issued from this
        # thread in sub-batch order on every rank.  The collective is `bss_gather_outputs` (C ABI: one NCCL group of
        # broadcasts on the sub-batch handle's own stream, straight into the final places) over a communicator of our own;
        # torch.distributed's all_gather serves when that communicator cannot be created.
        Bl = hi - lo
        esize = 8 if tdtype == torch.float64 else 4
        row_bytes = C * n_out * esize
        mode = os.environ.get('BSSGPU_GATHER_MODE', 'push')
        peers = self._peer_buffers(rank, world, group, B * row_bytes) if mode == 'push' else None
        if peers is not None:
            # peer-memory form: the result buffers of all ranks are mapped into every process (CUDA IPC) and each rank pushes
            # a finished sub-batch straight into its peers' buffers with device-to-device copies over NVLink -- copy engines,
            # so the update loops of the later sub-batches keep every SM (an NCCL kernel beside them does not overlap:
            # profiles/round2_scaling.md).  The buffers are reused by the next job: a barrier on entry keeps a fast rank
            # from writing into a buffer its owner is still reading, one on exit says that every push has landed.
            self.gather_backend = 'bss_push_outputs'
            dist.barrier(group=group)
            y_all = peers.as_tensor((B, C, n_out), tdtype)
            y_local = y_all[lo:hi]

            def push_part(i, plo, phi):
                self._parts[i][1].push_outputs(peers, y_local[plo:phi].data_ptr(), (lo + plo) * row_bytes, (phi - plo) * row_bytes)

            self.separate_waveform_batch(xl, fft_size, hop_size, window_fn, iteration=iteration,
                                         basis=None if basis is None else basis[lo:hi],
                                         activation=None if activation is None else activation[lo:hi], pipeline=pipeline,
                                         device_out=y_local.data_ptr(), loss_out=loss_out if loss_out is not None else np.zeros(Bl),
                                         on_done=push_part)
            for slot in self._parts:    # the pushes were queued on the sub-batch streams
                if slot is not None:
                    slot[1].synchronize()
            dist.barrier(group=group)
            return y_all.clone() if copy else y_all
        y_all = torch.empty((B, C, n_out), dtype=tdtype, device=device)
        y_local = y_all[lo:hi]
        comm = self._own_comm(rank, world, group)
        self.gather_backend = 'bss_gather_outputs' if comm is not None else 'torch.distributed.all_gather'

        def gather_part(i, plo, phi):
            if not overlap_gather:
                return
            if comm is not None:
                h = self._parts[i][1]
                h.gather_outputs(comm, y_local[plo:phi].data_ptr(), y_all.data_ptr() + plo * row_bytes, (phi - plo) * row_bytes,
                                 Bl * row_bytes)
            else:
                views = [y_all[r * Bl + plo:r * Bl + phi] for r in range(world)]
                dist.all_gather(views, y_local[plo:phi], group=group)

        self.separate_waveform_batch(xl, fft_size, hop_size, window_fn, iteration=iteration,
                                     basis=None if basis is None else basis[lo:hi],
                                     activation=None if activation is None else activation[lo:hi], pipeline=pipeline,
                                     device_out=y_local.data_ptr(), loss_out=loss_out if loss_out is not None else np.zeros(Bl),
                                     on_done=gather_part)
        if not overlap_gather:     # one collective after the last sub-batch (measurement aid: nothing overlaps it)
            if comm is not None:
                self._parts[0][1].gather_outputs(comm, y_local.data_ptr(), y_all.data_ptr(), Bl * row_bytes, Bl * row_bytes)
            else:
                dist.all_gather_into_tensor(y_all, y_local.clone(), group=group)
        if comm is not None:
            for slot in self._parts:    # the gathers were queued on the sub-batch streams
                if slot is not None:
                    slot[1].synchronize()
        return y_all

    def _peer_buffers(self, rank, world, group, nbytes):
        """The job's result buffers for the peer-memory gather, kept for the model's lifetime and grown on demand (None when
        CUDA IPC cannot be set up between the ranks; every rank then agrees on the NCCL gather through one small all-reduce)."""
        import torch
        import torch.distributed as dist
        cached = getattr(self, '_peers', None)
        if cached is not None and cached[0] == (rank, world) and (cached[1] is None or cached[1].nbytes >= nbytes):
            return cached[1]
        if cached is not None and cached[1] is not None:
            torch.cuda.synchronize(self.device)
            dist.barrier(group=group)     # nobody is still pushing into the buffers about to be released
            cached[1].unmap()
            dist.barrier(group=group)     # ... and nobody still maps the buffer its owner frees
            cached[1].close()
        peers = None
        self.gather_backend_error = None
        try:
            peers = _lib.PeerBuffers(rank, world, self.device, nbytes, group=group)
        except Exception as exc:   # reported through gather_backend / gather_backend_error; the NCCL gather serves
            self.gather_backend_error = "{}: {}".format(type(exc).__name__, exc)
        ok = torch.tensor([1 if peers is not None else 0], device=_flag_device(self.device, group))
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
        if int(ok.item()) == 0:
            if peers is not None:
                peers.close()
            peers = None
        self._peers = ((rank, world), peers)
        return peers

    def close(self, group=None):
        """Release what the whole-job calls keep between jobs: the sub-batch handles, and -- collectively, every rank calls it
        -- the peer-mapped result buffers and the communicator of the sharded job.  Tensors returned by the peer-memory form
        of `separate_waveform_batch_sharded` are invalid afterwards."""
        import torch.distributed as dist
        if self.handle is not None:
            self.handle.close()
            self.handle, self._key = None, None
        for slot in getattr(self, '_parts', []):
            if slot is not None:
                slot[1].close()
        self._parts = []
        peers = getattr(self, '_peers', None)
        if peers is not None and peers[1] is not None:
            dist.barrier(group=group)      # nobody still pushes into the buffers
            peers[1].unmap()
            dist.barrier(group=group)      # nobody still maps the buffer its owner frees
            peers[1].close()
        self._peers = None
        comm = getattr(self, '_comm', None)
        if comm is not None and comm[1] is not None:
            comm[1].close()
        self._comm = None

    def _own_comm(self, rank, world, group):
        """NCCL communicator for `bss_gather_outputs`, created once per model (None when NCCL cannot be set up that way; every
        rank then agrees on the fallback through one small all-reduce)."""
        import torch
        import torch.distributed as dist
        cached = getattr(self, '_comm', None)
        if cached is not None and cached[0] == (rank, world):
            return cached[1]
        comm = None
        earlier = getattr(self, 'gather_backend_error', None)     # why the peer-memory form was not taken, if it was tried
        try:
            if str(dist.get_backend(group)) == 'gloo':
                raise RuntimeError("gloo group: no NCCL communicator is made beside it")
            comm = _lib.NcclComm(rank, world, self.device, group=group)
        except Exception as exc:   # reported through gather_backend / gather_backend_error; the torch collective serves
            self.gather_backend_error = "{}{}: {}".format(earlier + " | " if earlier else "", type(exc).__name__, exc)
            comm = None
        ok = torch.tensor([1 if comm is not None else 0], device=_flag_device(self.device, group))
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
        if int(ok.item()) == 0:
            if comm is not None:
                comm.close()
            comm = None
        self._comm = ((rank, world), comm)
        return comm

    def separate_batch_sharded(self, X, iteration=100, basis=None, activation=None, group=None, pipeline='ramp', local_only=False):
        """The multi-GPU whole job (one process per GPU, torch.distributed initialised by the caller): every rank passes the
        SAME global batch description -- X (B,C,F,T) in host memory, of which it only reads its own contiguous shard
        `shard_range(B, rank, world)` -- runs its mixtures with `separate_batch` (copies pipelined against the update loop),
        leaves the estimates on its GPU and takes part in the one collective of the path, the NCCL all-gather of the
        separated outputs.  Returns a torch tensor (B,N,F,T) complex64 on this rank's GPU holding the estimates of ALL
        mixtures in batch order (`local_only=True`: only this rank's shard, no collective).  `basis` / `activation` are
        global (B,...) presets; when None every rank draws its own shard from NumPy's global state."""
        import torch
        import torch.distributed as dist
        world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        rank = dist.get_rank(group) if world > 1 else 0
        B, C, F, T = X.shape
        lo, hi = shard_range(B, rank, world)
        _check_presets(B, C, F, T, self.n_basis, None, basis, activation)
        device = torch.device('cuda', self.device)
        y_local = torch.empty((hi - lo, C, F, T), dtype=torch.complex64, device=device)
        if hi > lo:
            self.separate_batch(X[lo:hi], iteration=iteration, basis=None if basis is None else basis[lo:hi],
                                activation=None if activation is None else activation[lo:hi], pipeline=pipeline,
                                device_out=y_local.data_ptr())
        if local_only or world == 1:
            return y_local
        sizes = [shard_range(B, r, world)[1] - shard_range(B, r, world)[0] for r in range(world)]
        return torch.view_as_complex(gather_outputs(torch.view_as_real(y_local), world, group=group, sizes=sizes))

    def separate_waveforms(self, x, fft_size, hop_size=None, window_fn='hann', iteration=100, basis=None, activation=None,
                           dtype=np.float64):
        """Time domain in, time domain out: x (B,C,n_samples) real -> separated signals (B,N,n_out), n_out = the length
        scipy.signal.istft returns (>= n_samples).  STFT, update loop and ISTFT all run on the device: only waveforms cross
        PCIe (half the bytes of the spectrograms at 50 % overlap)."""
        from scipy import signal as ss
        x = np.ascontiguousarray(x, dtype=np.float32 if x.dtype == np.float32 else np.float64)
        B, C, n_samples = x.shape
        if hop_size is None:
            hop_size = fft_size // 2
        window = np.asarray(ss.get_window(window_fn, fft_size), dtype=np.float64)
        F, T = fft_size // 2 + 1, _lib.stft_frames(n_samples, fft_size, hop_size)
        K = self.n_basis
        _check_presets(B, C, F, T, K, None, basis, activation)
        h = self.open(B, C, F, T)
        h.reset_spatial()
        h.set_state(_lib.STATE_BASIS, np.random.rand(B, C, F, K) if basis is None else basis, np.float64)
        h.set_state(_lib.STATE_ACTIVATION, np.random.rand(B, C, K, T) if activation is None else activation, np.float64)
        h.set_input_waveform(x, fft_size, hop_size, window)
        h.run(iteration)
        return h.separate_waveform((B, C), fft_size, hop_size, window, dtype=dtype, projection_back=True)

    def update_once(self):
        self.handle.update_once()

    def compute_negative_loglikelihood(self):
        return self.handle.loss()

    @property
    def demix_filter(self):
        B, C, F, T = self.shape
        return self.handle.get_state(_lib.STATE_DEMIX_FILTER, (B, F, C, C), np.complex128)

    @property
    def basis(self):
        B, C, F, T = self.shape
        return self.handle.get_state(_lib.STATE_BASIS, (B, C, F, self.n_basis), np.float64)

    @property
    def activation(self):
        B, C, F, T = self.shape
        return self.handle.get_state(_lib.STATE_ACTIVATION, (B, C, self.n_basis, T), np.float64)


def ramp_sizes(n_items):
    """Sub-batch sizes of the pipelined whole-job call: B/16, B/8, 3B/16, B/4, 3B/16, B/8, B/16.  The upload of the first and
    the download of the last sub-batch are the only copies that nothing overlaps, so those are small; the middle ones are
    large because big launches are more efficient (64 mixtures: 4, 8, 12, 16, 12, 8, 4 -- measured 182 ms per job against
    192 ms for four equal parts, profiles/r5d_*).  Small batches fall back to at most four equal parts."""
    if n_items < 32:
        parts = max(1, min(4, n_items))
        return [hi - lo for lo, hi in (shard_range(n_items, i, parts) for i in range(parts))]
    weights = [1, 2, 3, 4, 3, 2, 1]
    sizes = [n_items * w // 16 for w in weights]
    sizes[3] += n_items - sum(sizes)
    return sizes


def shard_range(n_items, rank, world_size):
    """Contiguous shard [lo, hi) of rank `rank` (earlier ranks take the remainder)."""
    base, extra = divmod(n_items, world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def gather_outputs(local, world_size, group=None, sizes=None):
    """All-gather per-rank outputs (torch tensors, complex64 viewed as float32 pairs) along the batch axis, rank order ==
    batch order.  The single collective of the sharded path; runs on NCCL for CUDA tensors, gloo for CPU ones.
    `sizes`: per-rank batch sizes when they differ (`shard_range` gives the first ranks one more mixture when the batch is
    not a multiple of the world size); the shards are then padded to the largest one for the collective and trimmed after.
    With sizes=None every rank must hold the same number of mixtures (checked)."""
    import torch
    import torch.distributed as dist
    if world_size == 1:
        return local
    local = local.contiguous()
    if sizes is None:
        n = torch.tensor([local.shape[0]], dtype=torch.int64, device=local.device)
        counts = [torch.empty_like(n) for _ in range(world_size)]
        dist.all_gather(counts, n, group=group)
        sizes = [int(c.item()) for c in counts]
    sizes = [int(v) for v in sizes]
    if len(sizes) != world_size:
        raise ValueError("sizes has {} entries for {} ranks".format(len(sizes), world_size))
    rank = dist.get_rank(group)
    if sizes[rank] != local.shape[0]:
        raise ValueError("rank {} holds {} mixtures, sizes says {}".format(rank, local.shape[0], sizes[rank]))
    tail = tuple(local.shape[1:])
    if len(set(sizes)) == 1:
        out = torch.empty((world_size * sizes[0],) + tail, dtype=local.dtype, device=local.device)
        try:
            # one collective straight into the final buffer
            dist.all_gather_into_tensor(out, local, group=group)
        except (RuntimeError, NotImplementedError):
            dist.all_gather(list(out.chunk(world_size, dim=0)), local, group=group)
        return out
    big = max(sizes)
    padded = local
    if local.shape[0] < big:
        padded = torch.zeros((big,) + tail, dtype=local.dtype, device=local.device)
        padded[:local.shape[0]] = local
    parts = [torch.empty((big,) + tail, dtype=local.dtype, device=local.device) for _ in range(world_size)]
    dist.all_gather(parts, padded, group=group)
    return torch.cat([p[:n] for p, n in zip(parts, sizes)], dim=0)
