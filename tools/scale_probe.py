#!/usr/bin/env python
"""The sharded waveform job alone (BASELINE configs[4] as one call) under torchrun: median seconds of three jobs, max over ranks.
   Variants through the environment: BSSGPU_GATHER_MODE=push|nccl, BSSGPU_GATHER=allgather, BSSGPU_NCCL_MAX_CTAS=n,
   PROBE_OVERLAP=0 (one NCCL gather at the end)."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
from audio_source_separation_b200.batch import BatchedGaussILRMA, ramp_sizes

rank, world, local = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1)), int(os.environ.get('LOCAL_RANK', 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
B, C, F, T, K, FFT, HOP, steps = 64, 4, 2049, 512, 2, 4096, 2048, 20
n = (T - 1) * HOP
x = torch.empty((B, C, n), dtype=torch.int16, pin_memory=True)
rng = np.random.default_rng(rank)
x.numpy()[:] = rng.integers(-8000, 8000, size=(B, C, n), dtype=np.int16)
xn = x.numpy()


class G:
    shape = (world * B, C, n)

    def __getitem__(self, sl):
        return xn


T0 = np.broadcast_to(rng.random((C, F, K)), (world * B, C, F, K))
V0 = np.broadcast_to(rng.random((C, K, T)), (world * B, C, K, T))
m = BatchedGaussILRMA(n_basis=K, device=local)
loss = np.zeros(B)
overlap = os.environ.get('PROBE_OVERLAP', '1') != '0'
runs = []
for rep in range(4):
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    y = m.separate_waveform_batch_sharded(G(), FFT, HOP, iteration=steps, basis=T0, activation=V0, pipeline=ramp_sizes(B), loss_out=loss,
                                          overlap_gather=overlap)
    torch.cuda.synchronize()
    runs.append(time.perf_counter() - t0)
    del y
sec = float(np.median(runs[1:]))
t = torch.tensor([sec], device='cuda', dtype=torch.float64)
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    print(json.dumps({"world": world, "seconds": round(float(t.item()), 5), "it_per_s": round(world * B * steps / float(t.item())),
                      "mode": os.environ.get('BSSGPU_GATHER_MODE', 'push'), "gather": os.environ.get('BSSGPU_GATHER', 'bcast'), "max_ctas": os.environ.get('BSSGPU_NCCL_MAX_CTAS'),
                      "overlap": overlap, "backend": getattr(m, 'gather_backend', None), "timeline_rank0": m.timeline}))
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
