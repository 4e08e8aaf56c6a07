#!/usr/bin/env python
"""bench.py -- ILRMA iterations/sec (4ch x 2049 bins x 512 frames, K=2) on N GPUs vs the CPU reference path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

Workload: BASELINE.json configs[4] sharded -- every GPU holds B = 64 independent Gauss-ILRMA-IP problems of the
headline shape (configs[2]) resident in HBM (2.15 GB of complex64 STFT per GPU, far larger than L2, so every timed
iteration streams its inputs from DRAM; no L2 flush is needed).  One step = one update_once over the whole
resident batch; `value` = mixture-iterations per second = N * B * K / max-over-ranks device time.

  value      device-resident loop, CUDA events on the handle's stream, barrier + sync on both sides
  e2e        the same job through the public host API (BatchedGaussILRMA.__call__): H2D of the batch from pinned
             host memory, K iterations, separation + projection back, D2H of the result -- all timed
  roofline   the covariance-accumulate kernel timed alone (CUDA events) against the measured HBM peak
  cpu_baseline / --impl reference
             the oracle port of the reference's NumPy update (oracle/ilrma.py) on the host cores, one process
             per mixture, on a bounded sample of the same workload
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
GRAPH_PRIME = 10   # untimed iterations after the warm-up that capture the loop's CUDA graph (see run_gpu_arm)
METRIC = "ILRMA iterations/sec (4ch x 2049bin x 512frame, K=2), mixture-iterations summed over the batch"


# --------------------------------------------------------------------------------------------- inputs
def synth_batch(B, seed0, out=None):
    """Throughput input: complex Gaussian mixtures with a low-rank variance (cheap to generate for 64+ mixtures);
    parity on `mix2` inputs is covered by tests/."""
    if out is None:
        out = np.empty((B, C, F, T), dtype=np.complex64)
    for b in range(B):
        rng = np.random.default_rng(seed0 + b)
        Tb = 0.05 + rng.random((C, F, 1), dtype=np.float32)
        Vb = 0.05 + rng.random((C, 1, T), dtype=np.float32) ** 2
        S = np.sqrt(Tb * Vb * 0.5) * (rng.standard_normal((C, F, T), dtype=np.float32)
                                      + 1j * rng.standard_normal((C, F, T), dtype=np.float32))
        A = np.eye(C, dtype=np.complex64) + 0.35 * (rng.standard_normal((F, C, C), dtype=np.float32)
                                                    + 1j * rng.standard_normal((F, C, C), dtype=np.float32))
        out[b] = np.einsum('fij,jft->ift', A, S.astype(np.complex64))
        out[b] += 0.03 * (rng.standard_normal((C, F, T), dtype=np.float32) + 1j * rng.standard_normal((C, F, T), dtype=np.float32))
    return out


# --------------------------------------------------------------------------------------------- CPU arm
def _cpu_worker(args):
    seed, steps, warmup, barrier = args
    try:
        from threadpoolctl import threadpool_limits
        limiter = threadpool_limits(limits=1)
    except Exception:
        limiter = None
    from oracle import ilrma as o_ilrma, synth
    X = synth.mix2(C, F, T, seed=seed)
    W0, T0, V0 = synth.initial_state(C, F, T, K_BASIS, seed=7)
    st = o_ilrma.init_state(X, K_BASIS, W=W0, T=T0, V=V0)
    for _ in range(warmup):
        o_ilrma.update_once(st)
    barrier.wait()
    t0 = time.perf_counter()
    for _ in range(steps):
        o_ilrma.update_once(st)
    t1 = time.perf_counter()
    barrier.wait()
    del limiter
    return t0, t1


def cpu_reference_rate(steps, warmup, max_workers=None):
    """Mixture-iterations/sec of the oracle port with one process per mixture on the host cores."""
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
        spans = pool.map(_cpu_worker, [(1000 + w, steps, warmup, barrier) for w in range(workers)])
    t0 = min(s[0] for s in spans)
    t1 = max(s[1] for s in spans)
    elapsed = t1 - t0
    return workers * steps / elapsed, workers, elapsed


def run_reference_arm(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    steps = max(1, args.steps)
    warmup = max(0, args.warmup)
    # bounded sample: every worker runs at most `cap` update_once calls of ONE mixture (~1.5 s each on one core)
    cap = min(steps, 40)
    warm = min(warmup, 1)
    rate, workers, elapsed = cpu_reference_rate(cap, warm)
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": "iterations/s", "n_gpus": args.gpus, "steps": steps,
        "warmup": warmup, "ms_per_step": 1e3 * elapsed / cap, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": "Gauss-ILRMA-IP 4ch x 2049 bins x 512 frames, K=2, power normalisation (BASELINE configs[2]/[4])",
                   "timed_steps_per_worker": cap, "timed_warmup_per_worker": warm},
        "cpu_baseline": {"value": rate, "unit": "iterations/s", "cores": workers, "kind": "port",
                         "sample": "{} processes x {} update_once of one mix2(4,2049,512) mixture each (oracle/ilrma.py, NumPy "
                                   "float64, 1 BLAS thread per process), {} host cores present".format(workers, cap, os.cpu_count())},
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


# --------------------------------------------------------------------------------------------- GPU arm
def run_gpu_arm(args):
    import torch
    from audio_source_separation_b200 import _lib
    from audio_source_separation_b200.batch import BatchedGaussILRMA, gather_outputs

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

    # inputs in pinned host memory (what a caller hands to the public API)
    x_host = torch.empty((B, C, F, T), dtype=torch.complex64, pin_memory=True)
    synth_batch(B, 10_000 + rank * B, out=x_host.numpy())
    y_host = torch.empty((B, C, F, T), dtype=torch.complex64, pin_memory=True)
    rng = np.random.default_rng(7)
    T0 = rng.random((B, C, F, K_BASIS))
    V0 = rng.random((B, C, K_BASIS, T))

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

    # ---- covariance kernel alone -----------------------------------------------------------------
    cov_ms = h.time_covariance(20)
    cov_bytes = B * (8 * C * F * T + 8 * C * F * C * C + 4 * C * K_BASIS * (F + T))
    peak, peak_src = measured_hbm_peak()
    achieved = cov_bytes / (cov_ms * 1e-3) / 1e9
    step_bytes = B * (3 * 8 * C * F * T)
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
    x_np, y_np = x_host.numpy(), y_host.numpy()

    if args.pipeline == 'ramp':
        from audio_source_separation_b200.batch import ramp_sizes
        pipeline = ramp_sizes(B)
    else:
        pipeline = [int(v) for v in str(args.pipeline).split(',')]
        pipeline = pipeline[0] if len(pipeline) == 1 else pipeline

    def e2e_job():
        # the public whole-job call: 4 sub-batches on 4 streams so that H2D / D2H overlap the update loop
        model.separate_batch(x_np, y_np, iteration=steps, basis=T0, activation=V0, pipeline=pipeline)

    e2e_job()   # warm (allocations of the sub-batch handles and staging buffers)
    e2e_runs = []
    for _ in range(3):   # three whole jobs, the median is reported (each is ~0.2 s; a single one is noisy)
        barrier()
        t0 = time.perf_counter()
        e2e_job()
        e2e_runs.append(time.perf_counter() - t0)
    e2e_s = float(np.median(e2e_runs))
    if dist is not None:
        t = torch.tensor([e2e_s], device='cuda', dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_value = world * B * steps / e2e_s
    if not np.all(np.isfinite(y_host.numpy()[0, :, ::97, ::31])):
        raise RuntimeError("non-finite separated output")

    # ---- gather of the separated outputs (the only collective of the sharded path) ---------------
    gather_ms = None
    if dist is not None:
        y_dev = torch.empty((B, C, F, T), dtype=torch.complex64, device='cuda')
        h.separate_device(y_dev.data_ptr(), projection_back=True)
        h.synchronize()
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        gathered = gather_outputs(torch.view_as_real(y_dev), world)
        e1.record()
        torch.cuda.synchronize()
        gather_ms = e0.elapsed_time(e1)
        assert gathered.shape[0] == world * B
        del gathered, y_dev

    clocks = sampler.stop() if rank == 0 else None

    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            rate, workers, elapsed = cpu_reference_rate(4, 1)
            cpu = {"value": rate, "unit": "iterations/s", "cores": workers, "kind": "port",
                   "sample": "{} processes x 4 update_once of one mix2(4,2049,512) mixture each (oracle/ilrma.py, NumPy float64, "
                             "1 BLAS thread per process; {:.1f} s; {} host cores present)".format(workers, elapsed, os.cpu_count())}
        line = {
            "metric": METRIC, "value": value, "unit": "iterations/s", "n_gpus": world, "steps": steps, "warmup": warmup,
            "ms_per_step": ms / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": "BASELINE configs[4] shard: {} independent Gauss-ILRMA-IP mixtures per GPU, each 4ch x 2049 bins "
                                   "x 512 frames, K=2, power normalisation (configs[2] shape)".format(B),
                       "batch_per_gpu": B, "global_batch": world * B, "step": "one update_once over the resident batch",
                       "untimed": "{} warm-up + {} graph-priming iterations".format(warmup, GRAPH_PRIME),
                       "l2": "inputs larger than L2 ({:.2f} GB per GPU per pass): no flush".format(B * 8 * C * F * T / 1e9),
                       "storage": "complex64/float32 tensors, float64 per-bin solves",
                       "e2e_job": "one BatchedGaussILRMA.separate_batch call (pipelined sub-batches: {}): H2D batch from pinned memory "
                                  "+ {} iterations + separate/projection-back + D2H to pinned memory; bytes amortised per "
                                  "iteration".format(pipeline, steps),
                       "gather_ms": gather_ms},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "iterations/s", "h2d_bytes_per_step": x_host.numel() * 8 / steps,
                    "d2h_bytes_per_step": y_host.numel() * 8 / steps, "seconds": e2e_s,
                    "seconds_per_job": [round(v, 6) for v in e2e_runs]},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                         "kernel": "cov_kernel<C=4,NS=4,WM_ILRMA,K=2,CACHE>", "launch_ms": cov_ms, "algorithmic_bytes": cov_bytes,
                         "peak_source": peak_src},
            "roofline_step": {"algorithmic_bytes": step_bytes, "achieved": step_bytes / (ms / steps * 1e-3) / 1e9, "unit": "GB/s",
                              "frac": step_bytes / (ms / steps * 1e-3) / 1e9 / peak,
                              "note": "3 passes over X per iteration (basis MU, activation MU, covariance)"},
            "cpu_baseline": cpu,
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
