#!/usr/bin/env python
"""Multi-GPU check of the sharded product call (one process per GPU under torchrun): the tensor every rank gets from
separate_waveform_batch_sharded (peer-memory pushes, or with BSSGPU_GATHER_MODE=nccl bss_gather_outputs over our own NCCL
communicator, sub-batch by sub-batch) must equal, bit for bit, the rank-ordered concatenation of single-handle runs gathered with
torch.distributed.all_gather_into_tensor.
CHECK_BACKEND=gloo: the same check with a gloo group, which also allows two processes that SHARE one GPU (LOCAL_RANK=0 for
both; CUDA IPC works between processes on the same device) -- how tests/test_gpu_peer_exchange.py runs the peer-memory form
on a one-GPU box; the reference is gathered through host memory then."""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
from audio_source_separation_b200.batch import BatchedGaussILRMA

rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local)
GLOO = os.environ.get('CHECK_BACKEND', 'nccl') == 'gloo'
if GLOO:
    dist.init_process_group('gloo')
else:
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
Bl, C, K, fft, hop = 10, 4, 2, 1024, 256
B = Bl * world
from audio_source_separation_b200 import _lib
model = BatchedGaussILRMA(n_basis=K, device=local)
F = fft // 2 + 1
out = {}
for n in (20000, 31000):     # the second job needs larger result buffers (the peer buffers are re-made)
    rng = np.random.default_rng(3)
    pcm = rng.integers(-15000, 15000, size=(B, C, n)).astype(np.int16)       # the same global batch on every rank
    T = _lib.stft_frames(n, fft, hop)
    T0, V0 = rng.random((B, C, F, K)), rng.random((B, C, K, T))
    loss = np.zeros(Bl)
    for pipeline in ([2, 3, 5], 1):
        y_all = model.separate_waveform_batch_sharded(pcm, fft, hop, iteration=12, basis=T0, activation=V0, pipeline=pipeline, loss_out=loss)
        torch.cuda.synchronize()
        ref_local = BatchedGaussILRMA(n_basis=K, device=local).separate_waveform_batch(
            pcm[rank * Bl:(rank + 1) * Bl], fft, hop, iteration=12, basis=T0[rank * Bl:(rank + 1) * Bl], activation=V0[rank * Bl:(rank + 1) * Bl], pipeline=1)
        if GLOO:
            parts = [None] * world
            dist.all_gather_object(parts, ref_local)
            ref = torch.from_numpy(np.concatenate(parts)).cuda()
        else:
            ref = torch.empty((B, C, ref_local.shape[-1]), dtype=torch.float32, device='cuda')
            dist.all_gather_into_tensor(ref, torch.from_numpy(ref_local).cuda())
        same = bool(torch.equal(y_all, ref))
        out["n{} pipeline {}".format(n, pipeline)] = {
            "equal": same, "backend": model.gather_backend, "backend_error": getattr(model, 'gather_backend_error', None),
            "max_abs_diff": float((y_all - ref).abs().max())}
        assert np.all(np.isfinite(loss))
flag = torch.tensor([1 if all(v["equal"] for v in out.values()) else 0], device='cpu' if GLOO else 'cuda')
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print(json.dumps({"world": world, "all_ranks_equal": bool(flag.item()), "rank0": out}))
dist.barrier()
dist.destroy_process_group()
