#!/usr/bin/env python
"""bench.py -- ILRMA iterations/sec (4ch x 2049 bins x 512 frames, K=2) on N GPUs vs the CPU reference path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

Workload: BASELINE.json configs[4] sharded -- every GPU holds B = 64 independent Gauss-ILRMA-IP problems of the
headline shape (configs[2]) resident in HBM (2.15 GB of complex64 STFT per GPU, far larger than L2, so every timed
iteration streams its inputs from DRAM; no L2 flush is needed).  One step = one update_once over the whole
resident batch; `value` = mixture-iterations per second = N * B * K / max-over-ranks device time.  Both arms are fed the
SAME synthetic mixtures: `mix2(4, 2049, 512, seed)` of SURVEY.md Appendix D, complex64-rounded.

  value        device-resident loop, CUDA events on the handle's stream, barrier + sync on both sides
  e2e          BASELINE configs[4] through the public host API, host<->device copies inside the timed region:
               BatchedGaussILRMA.separate_waveform_batch_sharded -- the rank's waveforms up from pinned host memory, STFT +
               K iterations + projection back + ISTFT on the device, the separated signals of all ranks gathered over
               NVLink onto every GPU (the one exchange of the path, peer-memory pushes; none at N = 1), the final losses down.
               `e2e_host_waveform` / `e2e_host_spectrogram` deliver the outputs to pinned host memory instead (waveforms or
               STFT tensors both ways): those are bound by the host link of the box, not by the GPUs
  roofline     the covariance-accumulate kernel timed alone (CUDA events) against the measured HBM peak
  parity       the timed batch against a single-mixture handle replaying the same iterations (bit exact), and one more
               update_once against the CPU arm's implementation from the same state
  configs      (N = 1) BASELINE configs[0..3] -- EUC-NMF, AuxIVA-IP, single-mixture ILRMA, FastMNMF -- each with ms per
               iteration, the roofline of its covariance kernel where one is defined, and the CPU figure beside it
  cpu_baseline / --impl reference
               the reference's own classes (byte-compiled into oracle/_ref by oracle/build_ref.py, behind the NumPy-1
               linalg.solve shim; kind "reference"), or the oracle port when those are absent (kind "port"), on the host
               cores, one process per mixture, on a bounded sample of the same workload
Nothing here reads /root/reference.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

C, F, T, K_BASIS = 4, 2049, 512, 2
FFT, HOP = 4096, 2048            # STFT geometry of the headline shape: 2049 bins, 512 frames <- 1 046 528 samples per channel
GRAPH_PRIME = 10   # untimed iterations after the warm-up that capture the loop's CUDA graph (see run_gpu_arm)
METRIC = "ILRMA iterations/sec (4ch x 2049bin x 512frame, K=2), mixture-iterations summed over the batch"
WORKLOAD = "Gauss-ILRMA-IP 4ch x 2049 bins x 512 frames, K=2, power normalisation (BASELINE configs[2]/[4]), mix2 inputs"


# --------------------------------------------------------------------------------------------- inputs (both arms)
def mix2(n_ch, n_bins, n_frames, K=2, seed=0, snr_db=30.0, lo=0.05):
    """a determined mixture of low-rank-variance Gaussian sources, 30 dB white noise, rounded to
    complex64 (what the GPU stores) and returned as complex128 (what the CPU arm computes in)."""
    rng = np.random.default_rng(seed)
    Tb = lo + rng.random((n_ch, n_bins, K))
    Vb = lo + rng.random((n_ch, K, n_frames)) ** 2
    R = Tb @ Vb
    S = np.sqrt(R / 2) * (rng.standard_normal((n_ch, n_bins, n_frames)) + 1j * rng.standard_normal((n_ch, n_bins, n_frames)))
    A = np.eye(n_ch) + 0.5 * (rng.standard_normal((n_bins, n_ch, n_ch)) + 1j * rng.standard_normal((n_bins, n_ch, n_ch))) / np.sqrt(2)
    X = (A @ S.transpose(1, 0, 2)).transpose(1, 0, 2)
    p = np.mean(np.abs(X) ** 2)
    Nz = np.sqrt(p * 10 ** (-snr_db / 10) / 2) * (rng.standard_normal((n_ch, n_bins, n_frames)) + 1j * rng.standard_normal((n_ch, n_bins, n_frames)))
    return (X + Nz).astype(np.complex64).astype(np.complex128)


def initial_state(n_src, n_bins, n_frames, K, seed=7):
    """T0 ~ U(0,1) (N,F,K), V0 ~ U(0,1) (N,K,T), float32-representable; W = I."""
    rng = np.random.default_rng(seed)
    T0 = rng.random((n_src, n_bins, K)).astype(np.float32).astype(np.float64)
    V0 = rng.random((n_src, K, n_frames)).astype(np.float32).astype(np.float64)
    return T0, V0


def mixture_seed(global_index):
    return 1000 + int(global_index)


def mix2_batch(out, first_global_index):
    """out (B,C,F,T) complex64 <- mix2 mixtures first_global_index .. +B, generated on a few host threads."""
    from concurrent.futures import ThreadPoolExecutor

    def one(b):
        out[b] = mix2(C, F, T, seed=mixture_seed(first_global_index + b))

    workers = max(1, min(16, (os.cpu_count() or 2) // max(1, int(os.environ.get('LOCAL_WORLD_SIZE', os.environ.get('WORLD_SIZE', '1'))))))
    with ThreadPoolExecutor(max_workers=workers) as pool:
        list(pool.map(one, range(out.shape[0])))
    return out


# --------------------------------------------------------------------------------------------- CPU arm
def load_cpu_impl():
    """The reference's own classes when oracle/_ref holds them (kind 'reference'), else None (the oracle port is used)."""
    try:
        from oracle import build_ref
        return build_ref.load()
    except Exception:
        return None


class CpuIlrma:
    """One mixture on the CPU: the reference's GaussILRMA driven through its public update_once(), or the oracle port."""

    def __init__(self, X, T0, V0, W0=None):
        self.ref = load_cpu_impl()
        n_ch, n_bins, _ = X.shape
        if W0 is None:
            W0 = np.tile(np.eye(n_ch, dtype=np.complex128), (n_bins, 1, 1))
        if self.ref is not None:
            self.kind = 'reference'
            self.model = self.ref['GaussILRMA'](n_basis=T0.shape[-1], recordable_loss=False)
            self.model.input = X
            self.model._reset(demix_filter=W0, basis=T0, activation=V0)
        else:
            from oracle import ilrma as o_ilrma
            self.kind = 'port'
            self.o = o_ilrma
            self.st = o_ilrma.init_state(X, T0.shape[-1], W=W0, T=T0, V=V0)

    def update_once(self):
        if self.ref is not None:
            self.model.update_once()
        else:
            self.o.update_once(self.st)

    def state(self):
        if self.ref is not None:
            return self.model.demix_filter, self.model.basis, self.model.activation
        return self.st['W'], self.st['T'], self.st['V']


def _cpu_worker(args):
    seed, steps, warmup, barrier = args
    try:
        from threadpoolctl import threadpool_limits
        limiter = threadpool_limits(limits=1)
    except Exception:
        limiter = None
    X = mix2(C, F, T, seed=seed)
    T0, V0 = initial_state(C, F, T, K_BASIS)
    m = CpuIlrma(X, T0, V0)
    for _ in range(warmup):
        m.update_once()
    barrier.wait()
    t0 = time.perf_counter()
    for _ in range(steps):
        m.update_once()
    t1 = time.perf_counter()
    barrier.wait()
    del limiter
    return t0, t1, m.kind


def cpu_reference_rate(steps, warmup, max_workers=None):
    """Mixture-iterations/sec of the CPU implementation with one process per mixture on the host cores."""
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    avail_gb = 0.0
    try:
        with open('/proc/meminfo') as fh:
            for line in fh:
                if line.startswith('MemAvailable'):
                    avail_gb = int(line.split()[1]) / 1e6
    except Exception:
        pass
    workers = cores
    if avail_gb > 0:
        workers = min(workers, max(1, int(avail_gb // 4)))   # ~3 GB peak per worker (the reference's (N,F,T,C,C) temporary)
    if max_workers:
        workers = min(workers, max_workers)
    workers = max(1, min(workers, 64))
    ctx = mp.get_context('fork')
    mgr = ctx.Manager()
    barrier = mgr.Barrier(workers)
    with ctx.Pool(workers) as pool:
        spans = pool.map(_cpu_worker, [(mixture_seed(w), steps, warmup, barrier) for w in range(workers)])
    t0 = min(s[0] for s in spans)
    t1 = max(s[1] for s in spans)
    elapsed = t1 - t0
    per_proc = float(np.median([s[1] - s[0] for s in spans])) / steps
    return {"rate": workers * steps / elapsed, "workers": workers, "elapsed": elapsed, "kind": spans[0][2], "s_per_iter_one_process": per_proc}


def _cpu_one_update(args):
    """Checker: ONE update_once of the CPU implementation from a given state (used by the parity block)."""
    seed, W, T_, V = args
    try:
        from threadpoolctl import threadpool_limits
        limiter = threadpool_limits(limits=4)
    except Exception:
        limiter = None
    m = CpuIlrma(mix2(C, F, T, seed=seed), T_, V, W0=W)
    m.update_once()
    W1, T1, V1 = m.state()
    del limiter
    return np.array(W1), np.array(T1), np.array(V1), m.kind


def _cpu_config_worker(which):
    """CPU figure of one of BASELINE configs[0], [1], [3] (bounded samples; see `sample` in the result)."""
    try:
        from threadpoolctl import threadpool_limits
        limiter = threadpool_limits(limits=int(os.environ.get('BENCH_CPU_BLAS_THREADS', '8')))
    except Exception:
        limiter = None
    ref = load_cpu_impl()
    kind = 'reference' if ref is not None else 'port'
    out = None
    if which == 'cfg1':
        Z = cfg1_target()
        n_it = 50
        if ref is not None:
            np.random.seed(111)
            m = ref['EUCNMF'](n_basis=4)
            m.target = Z
            m._reset()
            m.update_once()
            t0 = time.perf_counter()
            for _ in range(n_it):
                m.update_once()
            dt = time.perf_counter() - t0
        else:
            from oracle import nmf as o_nmf
            rng = np.random.default_rng(1)
            Tm, Vm = rng.random((257, 4)), rng.random((4, 128))
            Tm, Vm = o_nmf.euc_step(Z, Tm, Vm)
            t0 = time.perf_counter()
            for _ in range(n_it):
                Tm, Vm = o_nmf.euc_step(Z, Tm, Vm)
            dt = time.perf_counter() - t0
        out = {"ms_per_iter": 1e3 * dt / n_it, "kind": kind, "sample": "50 update_once, one process"}
    elif which == 'cfg2':
        X = mix2(2, 1025, 256, seed=0)
        n_it = 6
        if ref is not None:
            m = ref['AuxLaplaceIVA'](recordable_loss=False)
            m.input = X
            m._reset()
            step = m.update_once
        else:
            from oracle import auxiva as o_auxiva
            st = o_auxiva.init_state(X, 'IP', None)
            step = lambda: o_auxiva.update_once(st, 'laplace', 'IP')   # noqa: E731
        step()
        t0 = time.perf_counter()
        for _ in range(n_it):
            step()
        dt = time.perf_counter() - t0
        out = {"ms_per_iter": 1e3 * dt / n_it, "kind": kind, "sample": "6 update_once, one process"}
    elif which == 'cfg4':
        # one iteration of the full shape takes ~30 s and 7 GB: time a slice of the bins (every term of the update is a
        # per-bin or per-(bin, frame) expression, the cost is linear in the number of bins) and scale it up
        Fs = 256
        X = mix2(8, Fs, 1024, seed=0)
        T0, V0 = initial_state(8, Fs, 1024, 2)
        if ref is not None:
            m = ref['FastMultichannelISNMF'](n_basis=2, recordable_loss=False)
            m.input = X
            m._reset(basis=T0, activation=V0)
            step = m.update_once
        else:
            from oracle import fastmnmf as o_mnmf
            st = o_mnmf.init_state(X, 2, 8, W=T0, H=V0)
            step = lambda: o_mnmf.update_once(st)   # noqa: E731
        t0 = time.perf_counter()
        step()
        dt = time.perf_counter() - t0
        out = {"ms_per_iter": 1e3 * dt * 2049 / Fs, "kind": kind,
               "sample": "1 update_once on {} of 2049 bins ({:.1f} s), scaled by 2049/{} (cost linear in bins)".format(Fs, dt, Fs)}
    del limiter
    return out


def cfg1_target():
    rng = np.random.default_rng(0)
    Z = (rng.standard_normal((257, 128)) ** 2 + rng.standard_normal((257, 128)) ** 2) / 2
    return Z.astype(np.float32).astype(np.float64)


def run_reference_arm(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    steps = max(1, args.steps)
    warmup = max(0, args.warmup)
    # bounded sample: every worker runs at most `cap` update_once calls of ONE mixture (~1.5 s each on one core)
    cap = min(steps, 40)
    warm = min(warmup, 1)
    r = cpu_reference_rate(cap, warm)
    rate, workers, elapsed = r['rate'], r['workers'], r['elapsed']
    what = ("the reference's own GaussILRMA.update_once (byte-compiled from the unmodified sources, oracle/_ref)" if r['kind'] == 'reference'
            else "oracle/ilrma.py (the port; oracle/_ref absent)")
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": "iterations/s", "n_gpus": args.gpus, "steps": steps,
        "warmup": warmup, "ms_per_step": 1e3 * elapsed / cap, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "timed_steps_per_worker": cap, "timed_warmup_per_worker": warm},
        "cpu_baseline": {"value": rate, "unit": "iterations/s", "cores": workers, "kind": r['kind'],
                         "sample": "{} processes x {} update_once of one mix2(4,2049,512) mixture each ({}; NumPy float64, "
                                   "1 BLAS thread per process), {} host cores present".format(workers, cap, what, os.cpu_count()),
                         "s_per_iter_one_process": r['s_per_iter_one_process']},
        "e2e": {"value": rate, "unit": "iterations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


# --------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    """Samples SM clock and throttle reasons through NVML while the bench runs."""

    def __init__(self, index):
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self._stop = threading.Event()
        self._thread = None
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.dev = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.dev, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def _loop(self):
        nv = self.nv
        names = {
            getattr(nv, 'nvmlClocksThrottleReasonHwSlowdown', 0x8): 'hw_slowdown',
            getattr(nv, 'nvmlClocksThrottleReasonHwThermalSlowdown', 0x40): 'hw_thermal_slowdown',
            getattr(nv, 'nvmlClocksThrottleReasonSwThermalSlowdown', 0x20): 'sw_thermal_slowdown',
            getattr(nv, 'nvmlClocksThrottleReasonSwPowerCap', 0x4): 'sw_power_cap',
            getattr(nv, 'nvmlClocksThrottleReasonHwPowerBrakeSlowdown', 0x80): 'hw_power_brake',
        }
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.dev, nv.NVML_CLOCK_SM))
                bits = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.dev)
                for bit, name in names.items():
                    if bits & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.02)

    def start(self):
        if self.ok:
            self._thread = threading.Thread(target=self._loop, daemon=True)
            self._thread.start()

    def stop(self):
        if self._thread:
            self._stop.set()
            self._thread.join()
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": 0}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


def measured_hbm_peak():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    try:
        with open(path) as fh:
            return float(json.load(fh)['hbm_gbs']), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def rel(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


# --------------------------------------------------------------------------------------------- BASELINE configs[0..3]
def bench_configs(device, peak):
    """GPU side of the `configs` block: ms per iteration of the device loop (CUDA events on the handle's stream) and, where
    SURVEY section 8d defines algorithmic bytes, the covariance kernel alone against the measured HBM peak."""
    from audio_source_separation_b200 import _lib
    out = {}

    def timed_loop(h, n_iter, reps):
        h.run(n_iter)   # warm: allocations, kernel attributes, graph capture
        h.synchronize()
        best = None
        for _ in range(reps):
            h.timer_begin()
            h.run(n_iter)
            ms = h.timer_end()
            best = ms if best is None else min(best, ms)
        return best / n_iter

    def cov_roofline(h, bytes_, repeat, kernel):
        ms = h.time_covariance(repeat)
        ach = bytes_ / (ms * 1e-3) / 1e9
        return {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "launch_ms": ms,
                "algorithmic_bytes": bytes_, "kernel": kernel, "traffic": None}

    # configs[0]: EUC-NMF K=4, 257 x 128, 50 iterations
    h = _lib.Handle(method=_lib.NMF_EUC, n_batch=1, n_channels=1, n_sources=1, n_bins=257, n_frames=128, n_basis=4, device=device)
    rng = np.random.default_rng(1)
    h.set_state(_lib.STATE_TARGET, cfg1_target(), np.float64)
    h.set_state(_lib.STATE_BASIS, rng.random((257, 4)), np.float64)
    h.set_state(_lib.STATE_ACTIVATION, rng.random((4, 128)), np.float64)
    out['cfg1'] = {"what": "EUC-NMF K=4 on a 257 x 128 spectrogram, 50 iterations per call", "ms_per_iter": timed_loop(h, 50, 5),
                   "roofline": None, "note": "0.26 MB problem: launch / latency bound (one cluster launch runs the whole loop)"}
    h.close()

    # configs[1]: AuxLaplaceIVA-IP 2ch x 1025 x 256, 30 iterations
    h = _lib.Handle(method=_lib.AUX_LAPLACE_IVA, spatial=_lib.SPATIAL_IP, normalize=_lib.NORMALIZE_NONE, n_batch=1, n_channels=2,
                    n_sources=2, n_bins=1025, n_frames=256, n_basis=1, device=device)
    h.set_input(mix2(2, 1025, 256, seed=0)[np.newaxis])
    h.reset_spatial()
    b2 = 8 * 2 * 1025 * 256 + 8 * 2 * 1025 * 4 + 4 * 2 * 256
    out['cfg2'] = {"what": "AuxLaplaceIVA-IP 2ch x 1025 bins x 256 frames, 30 iterations per call", "ms_per_iter": timed_loop(h, 30, 5),
                   "roofline": cov_roofline(h, b2, 50, "cov_kernel<C=2,WM_FRAME>"),
                   "note": "4.27 MB per launch: launch-latency dominated, reported only (SURVEY section 8d)"}
    h.close()

    # configs[2]: one Gauss-ILRMA mixture of the headline shape, 100 iterations
    h = _lib.Handle(method=_lib.GAUSS_ILRMA, spatial=_lib.SPATIAL_IP, normalize=_lib.NORMALIZE_POWER, n_batch=1, n_channels=C,
                    n_sources=C, n_bins=F, n_frames=T, n_basis=K_BASIS, device=device)
    T0, V0 = initial_state(C, F, T, K_BASIS)
    h.set_input(mix2(C, F, T, seed=mixture_seed(0))[np.newaxis])
    h.reset_spatial()
    h.set_state(_lib.STATE_BASIS, T0[np.newaxis], np.float64)
    h.set_state(_lib.STATE_ACTIVATION, V0[np.newaxis], np.float64)
    b3 = 8 * C * F * T + 8 * C * F * C * C + 4 * C * K_BASIS * (F + T)
    ms3 = timed_loop(h, 100, 3)
    out['cfg3_single'] = {"what": "Gauss-ILRMA-IP 4ch x 2049 bins x 512 frames, K=2, ONE mixture, 100 iterations per call",
                          "ms_per_iter": ms3, "iterations_per_s": 1e3 / ms3,
                          "roofline": cov_roofline(h, b3, 50, "cov_kernel<C=4,NS=4,WM_ILRMA,K=2>"),
                          "note": "34.7 MB per launch, L2 resident: the DRAM-bound case is the 64-mixture shard of the headline line"}
    # default user path: loss recorded after every iteration (recordable_loss=True), reduced on the device
    h.run_record(100)
    t0 = time.perf_counter()
    h.run_record(100)
    out['cfg3_single']["ms_per_iter_with_loss"] = 1e3 * (time.perf_counter() - t0) / 100
    h.close()

    # configs[3]: FastMNMF 8ch x 2049 x 1024, K=2, N=8, 50 iterations
    M, T4 = 8, 1024
    h = _lib.Handle(method=_lib.FAST_MNMF, normalize=_lib.NORMALIZE_POWER, n_batch=1, n_channels=M, n_sources=M, n_bins=F, n_frames=T4,
                    n_basis=2, device=device)
    T0, V0 = initial_state(M, F, T4, 2)
    h.set_input(mix2(M, F, T4, seed=0)[np.newaxis])
    h.reset_spatial()
    h.set_state(_lib.STATE_BASIS, T0[np.newaxis], np.float64)
    h.set_state(_lib.STATE_ACTIVATION, V0[np.newaxis], np.float64)
    b4 = 8 * M * F * T4 + 8 * M * F * M * M + 4 * (M * 2 * (F + T4) + M * F * M)
    out['cfg4'] = {"what": "FastMNMF 8ch x 2049 bins x 1024 frames, K=2, N=8, 50 iterations per call", "ms_per_iter": timed_loop(h, 50, 3),
                   "roofline": cov_roofline(h, b4, 20, "cov8_kernel (all 8 weighted covariances of update_diagonalizer, inverse variances computed in the kernel)"),
                   "note": "143.4 MB algorithmic per launch; weights are not credited (SURVEY section 8d)"}
    h.close()
    return out


# --------------------------------------------------------------------------------------------- GPU arm
def run_gpu_arm(args):
    import torch
    from audio_source_separation_b200 import _lib
    from audio_source_separation_b200.batch import BatchedGaussILRMA, gather_outputs, ramp_sizes

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    # stdout carries exactly one JSON line: keep NCCL's version banner (NCCL_DEBUG=VERSION) out of it
    if os.environ.get('NCCL_DEBUG', '').upper() in ('', 'VERSION'):
        os.environ['NCCL_DEBUG'] = 'WARN'
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: there is no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group(backend='nccl', device_id=torch.device('cuda', local_rank))

    B = args.batch
    steps, warmup = max(1, args.steps), max(3, args.warmup)

    # inputs in pinned host memory (what a caller hands to the public API): the same mix2 mixtures the CPU arm runs
    x_host = torch.empty((B, C, F, T), dtype=torch.complex64, pin_memory=True)
    mix2_batch(x_host.numpy(), rank * B)
    y_host = torch.empty((B, C, F, T), dtype=torch.complex64, pin_memory=True)
    T0s, V0s = initial_state(C, F, T, K_BASIS)
    T0 = np.ascontiguousarray(np.broadcast_to(T0s, (B,) + T0s.shape))
    V0 = np.ascontiguousarray(np.broadcast_to(V0s, (B,) + V0s.shape))

    model = BatchedGaussILRMA(n_basis=K_BASIS, device=local_rank)
    h = model.open(B, C, F, T)

    def upload():
        h.set_input_ptr(x_host.data_ptr(), _lib.C64)
        h.reset_spatial()
        h.set_state(_lib.STATE_BASIS, T0, np.float64)
        h.set_state(_lib.STATE_ACTIVATION, V0, np.float64)

    def barrier():
        h.synchronize()
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()

    upload()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()

    # ---- device-resident loop ------------------------------------------------------------------
    h.run(warmup)
    # bss_run replays loops of 10+ iterations from a CUDA graph that the handle captures once and keeps (api.cu:
    # graph_signature).  Capture + instantiation are one-off host work: on a busy host they took ~30 ms in the middle of the
    # timed loop (profiles/r5p_*), so the graph is primed here with GRAPH_PRIME more untimed iterations (an even count: the
    # basis buffers swap every iteration and the timed call must start where the capture did).
    h.run(GRAPH_PRIME)
    barrier()
    launches0 = h.launch_count()
    h.timer_begin()
    h.run(steps)
    ms = h.timer_end()
    launches = h.launch_count() - launches0
    barrier()
    if dist is not None:
        t = torch.tensor([ms], device='cuda')
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = world * B * steps / (ms * 1e-3)
    ip_kernel = h.get_info(_lib.INFO_IP_KERNEL)
    act_chunks = h.get_info(_lib.INFO_ACT_CHUNKS)
    source_model = h.get_info(_lib.INFO_SOURCE_MODEL)   # 2: fused single-pass source model, 1: three passes

    # ---- parity of the timed state -------------------------------------------------------------------
    # (a) mixture 0 of the batch against a single-mixture handle that replays the same calls on the same kernels and the
    #     same reduction order: bit exact.  (b) one more update_once of the batch against ONE update_once of the CPU arm's
    #     implementation (the reference's own class when oracle/_ref is there) from the state the timed loop ended in.
    parity = None
    if rank == 0:
        def state_of(hh, nb):
            return (hh.get_state(_lib.STATE_DEMIX_FILTER, (nb, F, C, C), np.complex128), hh.get_state(_lib.STATE_BASIS, (nb, C, F, K_BASIS), np.float64),
                    hh.get_state(_lib.STATE_ACTIVATION, (nb, C, K_BASIS, T), np.float64))
        Wb, Tb, Vb = state_of(h, B)
        s = _lib.Handle(method=_lib.GAUSS_ILRMA, spatial=_lib.SPATIAL_IP, normalize=_lib.NORMALIZE_POWER, n_batch=1, n_channels=C, n_sources=C,
                        n_bins=F, n_frames=T, n_basis=K_BASIS, device=local_rank)
        s.set_option(_lib.OPT_IP_KERNEL, ip_kernel)
        s.set_option(_lib.OPT_ACT_CHUNKS, act_chunks)
        s.set_input(x_host.numpy()[0:1])
        s.reset_spatial()
        s.set_state(_lib.STATE_BASIS, T0[0:1], np.float64)
        s.set_state(_lib.STATE_ACTIVATION, V0[0:1], np.float64)
        s.run(warmup)
        s.run(GRAPH_PRIME)
        s.run(steps)
        Ws, Ts, Vs = state_of(s, 1)
        s.close()
        parity = {"iterations": warmup + GRAPH_PRIME + steps,
                  "batch_vs_single_mixture": {"rel_W": rel(Wb[0], Ws[0]), "rel_T": rel(Tb[0], Ts[0]), "rel_V": rel(Vb[0], Vs[0]),
                                              "bit_identical": bool(np.array_equal(Wb[0], Ws[0]) and np.array_equal(Tb[0], Ts[0])
                                                                    and np.array_equal(Vb[0], Vs[0]))},
                  "ip_kernel": {1: "ip_sweep_kernel (thread per bin)", 2: "ip_sweep_group_kernel", 3: "fused in cov_kernel"}.get(ip_kernel, ip_kernel),
                  "source_model": "fused single pass (mu_fused_kernel)" if source_model == _lib.SOURCE_MODEL_FUSED else "three passes",
                  "act_chunks": act_chunks, "state_before": (Wb[0], Tb[0], Vb[0])}
    h.update_once()          # one more (untimed) iteration on every rank: keeps the ranks in step, feeds check (b)
    if rank == 0:
        parity["state_after"] = (h.get_state(_lib.STATE_DEMIX_FILTER, (B, F, C, C), np.complex128)[0],
                                 h.get_state(_lib.STATE_BASIS, (B, C, F, K_BASIS), np.float64)[0],
                                 h.get_state(_lib.STATE_ACTIVATION, (B, C, K_BASIS, T), np.float64)[0])

    # ---- covariance kernel alone -----------------------------------------------------------------
    cov_ms = h.time_covariance(20)
    cov_bytes = B * (8 * C * F * T + 8 * C * F * C * C + 4 * C * K_BASIS * (F + T))
    peak, peak_src = measured_hbm_peak()
    achieved = cov_bytes / (cov_ms * 1e-3) / 1e9
    passes = 2 if source_model == _lib.SOURCE_MODEL_FUSED else 3
    step_bytes = B * (passes * 8 * C * F * T)
    traffic = None   # DRAM bytes of one launch from the committed ncu --set full capture (same workload only)
    try:
        with open(os.path.join(ROOT, 'profiles', 'cov_kernel_traffic.json')) as fh:
            tr = json.load(fh)
        if tr.get('batch_per_gpu') == B:
            traffic = tr['dram_bytes_read'] + tr['dram_bytes_write']
    except Exception:
        traffic = None
    loss = h.loss()
    if not np.all(np.isfinite(loss)):
        raise RuntimeError("non-finite loss after the timed loop")

    # ---- end to end through the host API ---------------------------------------------------------
    if args.pipeline == 'ramp':
        pipeline = ramp_sizes(B)
    else:
        pipeline = [int(v) for v in str(args.pipeline).split(',')]
        pipeline = pipeline[0] if len(pipeline) == 1 else pipeline

    def timed_jobs(job, n=3):
        job()   # warm (allocations of the sub-batch handles, staging buffers, graph capture)
        runs = []
        for _ in range(n):   # whole jobs, the median is reported (each is a fraction of a second; a single one is noisy)
            barrier()
            t0 = time.perf_counter()
            job()
            runs.append(time.perf_counter() - t0)
        sec = float(np.median(runs))
        if dist is not None:
            tt = torch.tensor([sec], device='cuda', dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            sec = float(tt.item())
        return sec, runs

    # (1) spectrograms across PCIe
    x_np, y_np = x_host.numpy(), y_host.numpy()
    spec_s, spec_runs = timed_jobs(lambda: model.separate_batch(x_np, y_np, iteration=steps, basis=T0, activation=V0, pipeline=pipeline))
    if not np.all(np.isfinite(y_np[0, :, ::97, ::31])):
        raise RuntimeError("non-finite separated output")
    # (2) waveforms across PCIe: the time-domain signals whose STFT has the headline shape (white noise through a random
    #     mixing matrix: throughput input; the STFT feed has its own parity tests)
    n_samples = (T - 1) * HOP
    # 16-bit PCM, the format recordings come in (the reference's notebooks read int16 wav files and divide by 32768)
    wave_in = torch.empty((B, C, n_samples), dtype=torch.int16, pin_memory=True)
    rng = np.random.default_rng(50_000 + rank)
    for b in range(B):
        src = rng.standard_normal((C, n_samples), dtype=np.float32) * rng.random((C, 1), dtype=np.float32)
        mix = (np.eye(C, dtype=np.float32) + 0.4 * rng.standard_normal((C, C), dtype=np.float32)) @ src
        wave_in.numpy()[b] = np.clip(mix * 4000.0, -32767, 32767).astype(np.int16)
    n_out = _lib.istft_length(T, FFT, HOP)
    wave_out = torch.empty((B, C, n_out), dtype=torch.float32, pin_memory=True)
    assert _lib.stft_frames(n_samples, FFT, HOP) == T
    w_in, w_out = wave_in.numpy(), wave_out.numpy()
    wave_model = BatchedGaussILRMA(n_basis=K_BASIS, device=local_rank)
    wave_s, wave_runs = timed_jobs(lambda: wave_model.separate_waveform_batch(w_in, FFT, HOP, out=w_out, iteration=steps, basis=T0,
                                                                             activation=V0, pipeline=pipeline))
    if not np.all(np.isfinite(w_out[0, :, ::997])):
        raise RuntimeError("non-finite separated waveform")
    timelines = {"spectrogram": getattr(model, 'timeline', None), "waveform": getattr(wave_model, 'timeline', None)}
    del wave_model

    # ---- BASELINE configs[4] as one product call: shard up, iterate, NVLink gather of the outputs ----------------------
    # every rank describes the same global batch; only its own shard is read (here: the rank's pinned buffer)
    class GlobalWaveforms:   # minimal array protocol: shape + slicing of the local shard
        shape = (world * B, C, n_samples)

        def __getitem__(self, sl):
            assert sl.start == rank * B and sl.stop == (rank + 1) * B
            return w_in
    gw = GlobalWaveforms()
    T0g = np.broadcast_to(T0s, (world * B,) + T0s.shape)
    V0g = np.broadcast_to(V0s, (world * B,) + V0s.shape)
    losses = np.zeros(B, dtype=np.float64)
    shard_model = BatchedGaussILRMA(n_basis=K_BASIS, device=local_rank)
    holder = {}

    def sharded_job():
        holder['y'] = None   # the previous job's gathered output is released before the next one allocates
        holder['y'] = shard_model.separate_waveform_batch_sharded(gw, FFT, HOP, iteration=steps, basis=T0g, activation=V0g,
                                                                  pipeline=pipeline, loss_out=losses)
        torch.cuda.synchronize()
    shard_s, shard_runs = timed_jobs(sharded_job)
    if tuple(holder['y'].shape) != (world * B, C, n_out) or not np.all(np.isfinite(losses)):
        raise RuntimeError("sharded job: wrong output shape or non-finite loss")
    timelines["sharded"] = getattr(shard_model, 'timeline', None)
    gather_backend = (getattr(shard_model, 'gather_backend', None), getattr(shard_model, 'gather_backend_error', None))
    holder.clear()
    shard_model.close()     # collective at N > 1: the peer-mapped result buffers go back before the gather microbenchmark
    del shard_model

    # the all-gather alone, warm: (world - 1) x the per-rank outputs received per GPU
    gather = None
    if dist is not None:
        gather = {"nvlink5_unidirectional_GBps": 900.0}
        for label, shape, dt in (("spectrogram_outputs", (B, C, F, T, 2), torch.float32), ("waveform_outputs", (B, C, n_out), torch.float32)):
            y_dev = torch.ones(shape, dtype=dt, device='cuda')
            gather_outputs(y_dev, world)   # warm: communicator channels, buffers
            torch.cuda.synchronize()
            times = []
            for _ in range(3):
                dist.barrier()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                gathered = gather_outputs(y_dev, world)
                e1.record()
                torch.cuda.synchronize()
                times.append(e0.elapsed_time(e1))
                assert gathered.shape[0] == world * B
                del gathered
            tt = torch.tensor([float(np.median(times))], device='cuda')
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            g_ms = float(tt.item())
            recv = (world - 1) * y_dev.numel() * 4
            gather[label] = {"ms": g_ms, "bytes_received_per_gpu": recv, "GBps_received_per_gpu": recv / (g_ms * 1e-3) / 1e9}
            del y_dev

    clocks = sampler.stop() if rank == 0 else None
    if os.environ.get('BENCH_TIMELINE'):   # per-rank phase marks of the last end-to-end jobs (diagnostics)
        with open('{}.rank{}.json'.format(os.environ['BENCH_TIMELINE'], rank), 'w') as fh:
            json.dump({"rank": rank, "world": world, "spectrogram_s": spec_s, "waveform_s": wave_s, "sharded_s": shard_s, "timelines": timelines,
                       "cpus": len(os.sched_getaffinity(0))}, fh)

    if rank == 0:
        cpu, cfgs = None, None
        if world == 1 and not args.no_cpu_baseline:
            import multiprocessing as mp
            r = cpu_reference_rate(4, 1)
            what = ("the reference's own GaussILRMA.update_once, byte-compiled from the unmodified sources (oracle/_ref)" if r['kind'] == 'reference'
                    else "oracle/ilrma.py (the port; oracle/_ref absent)")
            cpu = {"value": r['rate'], "unit": "iterations/s", "cores": r['workers'], "kind": r['kind'],
                   "sample": "{} processes x 4 update_once of one mix2(4,2049,512) mixture each ({}; NumPy float64, 1 BLAS thread per "
                             "process; {:.1f} s; {} host cores present)".format(r['workers'], what, r['elapsed'], os.cpu_count()),
                   "s_per_iter_one_process": r['s_per_iter_one_process']}
            # parity (b): the CPU implementation as the checker of one update from the timed state
            with mp.get_context('fork').Pool(1) as pool:
                W1, T1, V1, kind = pool.apply(_cpu_one_update, ((mixture_seed(0),) + tuple(parity["state_before"]),))
            Wg, Tg, Vg = parity["state_after"]
            parity["one_update_vs_cpu"] = {"rel_W": rel(Wg, W1), "rel_T": rel(Tg, T1), "rel_V": rel(Vg, V1), "cpu_kind": kind,
                                           "tolerance": 2e-4}
            # BASELINE configs[0..3] in the same clock-sampled run
            cfgs = bench_configs(local_rank, peak)
            with mp.get_context('fork').Pool(1) as pool:
                for name in ('cfg1', 'cfg2', 'cfg4'):
                    cfgs[name]["cpu"] = pool.apply(_cpu_config_worker, (name,))
            cfgs['cfg3_single']["cpu"] = {"ms_per_iter": 1e3 * r['s_per_iter_one_process'], "kind": r['kind'],
                                          "sample": "median over the {} concurrent single-threaded processes of cpu_baseline".format(r['workers'])}
            for name in cfgs:
                cfgs[name]["speedup_vs_cpu"] = cfgs[name]["cpu"]["ms_per_iter"] / cfgs[name]["ms_per_iter"]
        if parity is not None:
            parity.pop("state_before", None)
            parity.pop("state_after", None)
        wave_h2d = wave_in.numel() * 2
        wave_d2h = wave_out.numel() * 4
        line = {
            "metric": METRIC, "value": value, "unit": "iterations/s", "n_gpus": world, "steps": steps, "warmup": warmup,
            "ms_per_step": ms / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": "BASELINE configs[4] shard: {} independent Gauss-ILRMA-IP mixtures per GPU, each 4ch x 2049 bins "
                                   "x 512 frames, K=2, power normalisation (configs[2] shape), mix2 inputs".format(B),
                       "batch_per_gpu": B, "global_batch": world * B, "step": "one update_once over the resident batch",
                       "untimed": "{} warm-up + {} graph-priming iterations".format(warmup, GRAPH_PRIME),
                       "l2": "inputs larger than L2 ({:.2f} GB per GPU per pass): no flush".format(B * 8 * C * F * T / 1e9),
                       "storage": "complex64/float32 tensors, float64 per-bin solves",
                       "e2e_job": "BASELINE configs[4] as one call, BatchedGaussILRMA.separate_waveform_batch_sharded (pipelined "
                                  "sub-batches: {}): H2D of the rank's int16 PCM waveforms ({} samples x 4ch per mixture) from pinned memory + "
                                  "STFT ({}/{}) + {} iterations + separate/projection-back + ISTFT on the device + all-gather of all "
                                  "separated float32 signals onto every GPU over NVLink (pushed into the peers' result buffers through "
                                  "peer memory, bss_push_outputs; BSSGPU_GATHER_MODE=nccl: NCCL broadcasts), sub-batch by sub-batch "
                                  "behind the update loops of the later ones (N = 1: no exchange, the signals stay on the GPU) + D2H of the "
                                  "final per-mixture losses; bytes amortised per iteration.  e2e_host_* are the same job with the "
                                  "outputs delivered to pinned host memory instead (bound by the host link of the box, see "
                                  "profiles/r6d_pcie_probe_8gpu.json)".format(pipeline, n_samples, FFT, HOP, steps)},
            "clocks": clocks,
            "e2e": {"value": world * B * steps / shard_s, "unit": "iterations/s", "h2d_bytes_per_step": wave_h2d / steps,
                    "d2h_bytes_per_step": B * 8 / steps, "seconds": shard_s, "seconds_per_job": [round(v, 6) for v in shard_runs],
                    "job": "separate_waveform_batch_sharded: waveforms up, separated waveforms gathered over NVLink onto every GPU, "
                           "losses down",
                    "gather_backend": gather_backend[0] if world > 1 else None, "gather_backend_error": gather_backend[1]},
            "e2e_host_waveform": {"value": world * B * steps / wave_s, "unit": "iterations/s", "h2d_bytes_per_step": wave_h2d / steps,
                                  "d2h_bytes_per_step": wave_d2h / steps, "seconds": wave_s, "seconds_per_job": [round(v, 6) for v in wave_runs],
                                  "job": "separate_waveform_batch: waveforms up, separated waveforms down to pinned host memory"},
            "e2e_host_spectrogram": {"value": world * B * steps / spec_s, "unit": "iterations/s", "h2d_bytes_per_step": x_host.numel() * 8 / steps,
                                     "d2h_bytes_per_step": y_host.numel() * 8 / steps, "seconds": spec_s,
                                     "seconds_per_job": [round(v, 6) for v in spec_runs],
                                     "job": "separate_batch: complex64 STFT tensors up, separated STFT tensors down to pinned host memory"},
            "gather": gather,
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                         "kernel": "cov_kernel<C=4,NS=4,WM_ILRMA,K=2,CACHE>", "launch_ms": cov_ms, "algorithmic_bytes": cov_bytes,
                         "peak_source": peak_src},
            "roofline_step": {"algorithmic_bytes": step_bytes, "achieved": step_bytes / (ms / steps * 1e-3) / 1e9, "unit": "GB/s",
                              "frac": step_bytes / (ms / steps * 1e-3) / 1e9 / peak,
                              "passes_over_X": passes,
                              "note": ("2 passes over X per iteration (fused basis + activation update, covariance)" if passes == 2 else
                                       "3 passes over X per iteration (basis MU, activation MU, covariance)")},
            "parity": parity,
            "cpu_baseline": cpu,
            "configs": cfgs,
        }
        print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=100)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--batch', type=int, default=64, help='mixtures per GPU')
    ap.add_argument('--pipeline', default='ramp', help="sub-batches of the end-to-end job (copy/compute overlap): 'ramp' "
                    "(batch.ramp_sizes), a count, or comma-separated sub-batch sizes adding up to --batch")
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    args = ap.parse_args()
    if args.impl == 'reference':
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == '__main__':
    main()
