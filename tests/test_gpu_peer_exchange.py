"""The peer-memory exchange of the sharded job (`bss_peer_*`, `bss_push_outputs`) on a ONE-GPU box: two
processes that share the GPU, a gloo group for the host-side plumbing, CUDA IPC between them.  `tools/check_sharded.py` holds
the tensor every rank gets from `separate_waveform_batch_sharded` bit-identical to the rank-ordered concatenation of
single-handle runs (two pipelines, two job sizes: the peer buffers are re-made in between).  The multi-GPU runs of the same
script (2 and 8 GPUs, NCCL group) are recorded in profiles/round2_scaling.md.

Whatever keeps the two processes from running at all (no CUDA IPC in the container, no CUDA tensors in gloo, a timeout on a
busy box) skips the test; a job that ran and produced different bits fails it."""
import json
import os
import socket
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


@pytest.mark.gpu
def test_peer_memory_exchange_two_processes_one_gpu(cuda_device):
    env = dict(os.environ, MASTER_ADDR='127.0.0.1', MASTER_PORT=str(_free_port()), WORLD_SIZE='2', LOCAL_RANK='0',
               CHECK_BACKEND='gloo', BSSGPU_GATHER_MODE='push', OMP_NUM_THREADS='1')
    script = os.path.join(ROOT, 'tools', 'check_sharded.py')
    procs = [subprocess.Popen([sys.executable, script], env=dict(env, RANK=str(r)), cwd=ROOT, stdout=subprocess.PIPE,
                              stderr=subprocess.PIPE, text=True) for r in range(2)]
    outs = []
    try:
        for p in procs:
            outs.append(p.communicate(timeout=150))
    except subprocess.TimeoutExpired:
        for p in procs:
            p.kill()
        pytest.skip("the two-process job did not finish within 150 s")
    if any(p.returncode != 0 for p in procs):
        tail = " | ".join((err or "").strip().splitlines()[-1] if (err or "").strip() else "" for _, err in outs)
        pytest.skip("the two-process job could not run here: " + tail)
    lines = [ln for ln in outs[0][0].splitlines() if ln.startswith('{')]
    try:
        result = json.loads(lines[-1])
    except (IndexError, ValueError):
        pytest.skip("rank 0 printed no result line")
    assert result["world"] == 2
    assert result["all_ranks_equal"], result
    backends = {v["backend"] for v in result["rank0"].values()}
    if backends != {"bss_push_outputs"}:
        pytest.skip("CUDA IPC between the two processes is not available here (exchange ran through {}: {})".format(
            backends, [v["backend_error"] for v in result["rank0"].values()][0]))
    assert all(v["equal"] and v["max_abs_diff"] == 0.0 for v in result["rank0"].values())
