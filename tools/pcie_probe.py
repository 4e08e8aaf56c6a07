#!/usr/bin/env python
"""Host <-> device copy bandwidth of pinned buffers with 1..N GPUs active at once (one process per GPU under torchrun).

Answers one question for the end-to-end job: when every rank of a node uploads its shard and downloads its outputs at the
same time, is the aggregate limited by the host (memory / PCIe root / IOMMU of the VM) rather than by each GPU's own link?

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29631 tools/pcie_probe.py
"""
import json
import os
import time

import torch
import torch.distributed as dist

rank, world, local = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1)), int(os.environ.get('LOCAL_RANK', 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
GB = 1 << 30
n = GB // 4
h_in = torch.empty(n, dtype=torch.float32, pin_memory=True).fill_(1.0)
h_out = torch.empty(n, dtype=torch.float32, pin_memory=True)
d_in = torch.empty(n, dtype=torch.float32, device='cuda')
d_out = torch.ones(n, dtype=torch.float32, device='cuda')
s_up, s_dn = torch.cuda.Stream(), torch.cuda.Stream()


def barrier():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()


def run(up, down, active, reps=4):
    """GB/s of this rank (0 when it sits this round out); every rank takes part in the barriers."""
    best = 0.0
    for _ in range(reps):
        barrier()
        t0 = time.perf_counter()
        if active:
            if up:
                with torch.cuda.stream(s_up):
                    d_in.copy_(h_in, non_blocking=True)
            if down:
                with torch.cuda.stream(s_dn):
                    h_out.copy_(d_out, non_blocking=True)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            best = max(best, ((1 if up else 0) + (1 if down else 0)) * GB / dt / 1e9)
    return best


out = {}
for name, up, down in (('h2d', True, False), ('d2h', False, True), ('both', True, True)):
    for k in sorted({1, 2, 4, world} & set(range(1, world + 1))):
        mine = run(up, down, rank < k)
        t = torch.tensor([mine], device='cuda', dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t)
        out['{}_{}active_aggregate_GBps'.format(name, k)] = round(float(t.item()), 1)
if rank == 0:
    try:
        out['cpus'] = len(os.sched_getaffinity(0))
        with open('/proc/meminfo') as fh:
            out['mem_total_gb'] = round(int(fh.readline().split()[1]) / 1e6, 1)
    except Exception:
        pass
    print(json.dumps(out))
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
