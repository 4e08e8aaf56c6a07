"""CPU-only checks of the boundary and of the multi-GPU host logic:

  * libbssgpu.so loads and exports every symbol include/bssgpu.h declares (no compute call is made);
  * without a CUDA device the product fails loudly instead of falling back to a CPU path;
  * the product package never imports the oracle;
  * the batch sharding / final gather of the N > 1 path on world_size-2 `gloo` process groups.
"""
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, 'include', 'bssgpu.h')
PKG = os.path.join(ROOT, 'audio_source_separation_b200')


def _declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(bss_[a-z_0-9]+)\s*\(', text)))


@pytest.fixture(scope='module')
def lib():
    from audio_source_separation_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    return _lib.load()


def test_header_symbols_are_exported_and_bound(lib):
    from audio_source_separation_b200 import _lib
    names = _declared_symbols()
    assert len(names) >= 25
    for name in names:
        assert hasattr(lib, name), "libbssgpu.so does not export " + name
    # the ctypes table binds exactly the header's entry points
    assert sorted(_lib.SIGNATURES) == names
    assert lib.bss_version().decode().startswith('bssgpu')


def test_config_struct_matches_header(lib):
    """struct bss_config: 14 int32 followed by 4 doubles, no padding surprises."""
    import ctypes
    from audio_source_separation_b200 import _lib
    assert ctypes.sizeof(_lib.Config) == 14 * 4 + 4 * 8
    text = open(HEADER).read()
    body = text[text.index('typedef struct bss_config {'):text.index('} bss_config;')]
    fields = re.findall(r'\b(?:int32_t|double)\s+([a-z_]+);', body)
    assert fields == [f[0] for f in _lib.Config._fields_]


def test_no_cpu_fallback(lib):
    """Here (no GPU) creating a handle must fail with a CUDA error, not compute on the host."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from audio_source_separation_b200.bss.ilrma import GaussILRMA
    from audio_source_separation_b200.algorithm.nmf import EUCNMF
    X = (np.random.default_rng(0).standard_normal((2, 5, 8)) + 0j)
    with pytest.raises(RuntimeError, match='no CUDA device'):
        GaussILRMA(n_basis=2)(X, iteration=1)
    with pytest.raises(RuntimeError, match='no CUDA device'):
        EUCNMF(n_basis=2)(np.abs(X[0]), iteration=1)


def test_product_never_imports_oracle():
    for dirpath, _, files in os.walk(PKG):
        for fn in files:
            if fn.endswith(('.py', '.cu', '.cuh', '.h')):
                text = open(os.path.join(dirpath, fn)).read()
                assert not re.search(r'^\s*(from|import)\s+oracle\b', text, flags=re.M), fn
                assert '/root/reference' not in text, fn


def test_shard_range_partitions_the_batch():
    from audio_source_separation_b200.batch import shard_range
    for n_items in (0, 1, 7, 64, 512, 513):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n_items, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n_items
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    assert shard_range(512, 3, 8) == (192, 256)


_WORKER = r'''
import os, sys
sys.path.insert(0, {root!r})
import numpy as np, torch, torch.distributed as dist
from audio_source_separation_b200.batch import shard_range, gather_outputs
rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
dist.init_process_group(backend='gloo')
B, N, F, T = 6, 2, 5, 4
full = (np.arange(B * N * F * T * 2, dtype=np.float32)).reshape(B, N, F, T, 2)
lo, hi = shard_range(B, rank, world)
local = torch.from_numpy(full[lo:hi].copy())          # what this rank's GPU would have separated
out = gather_outputs(local, world)
assert out.shape == (B, N, F, T, 2), out.shape
assert np.array_equal(out.numpy(), full)               # rank order == batch order, bit exact
# a batch that does not divide by the world size: uneven shards, padded for the collective and trimmed after
B2 = 7
full2 = (np.arange(B2 * N * F * T * 2, dtype=np.float32) * 0.5).reshape(B2, N, F, T, 2)
lo, hi = shard_range(B2, rank, world)
local2 = torch.from_numpy(full2[lo:hi].copy())
sizes = [shard_range(B2, r, world)[1] - shard_range(B2, r, world)[0] for r in range(world)]
assert sizes == [4, 3]
for given in (sizes, None):
    out2 = gather_outputs(local2, world, sizes=given)
    assert out2.shape == (B2, N, F, T, 2) and np.array_equal(out2.numpy(), full2)
try:
    gather_outputs(local2, world, sizes=[3, 4])
    raise SystemExit('a wrong size table must be rejected')
except ValueError:
    pass
dist.barrier()
dist.destroy_process_group()
print('ok', rank)
'''


def test_gather_outputs_world_size_2_gloo(tmp_path):
    """The only collective of the sharded path (all-gather of the separated outputs) on 2 CPU ranks."""
    script = tmp_path / 'worker.py'
    script.write_text(_WORKER.format(root=ROOT))
    env = dict(os.environ, MASTER_ADDR='127.0.0.1')
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node=2', '--master-addr', '127.0.0.1',
           '--master-port', '29611', str(script)]
    res = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stdout + res.stderr
    assert res.stdout.count('ok') == 2


def test_batched_presets_are_shape_checked_before_the_library_sees_them():
    """The C ABI carries no lengths: wrongly shaped presets must raise in Python."""
    from audio_source_separation_b200.batch import _check_presets
    B, C, F, T, K = 3, 2, 5, 7, 2
    _check_presets(B, C, F, T, K, np.zeros((B, F, C, C)), np.zeros((B, C, F, K)), np.zeros((B, C, K, T)))
    _check_presets(B, C, F, T, K)
    for bad in (dict(basis=np.zeros((B, C, F, K + 1))), dict(activation=np.zeros((B, C, K, T - 1))),
                dict(demix_filter=np.zeros((F, C, C))), dict(basis=np.zeros((C, F, K)))):
        with pytest.raises(ValueError):
            _check_presets(B, C, F, T, K, **bad)


def test_ramp_sizes_cover_the_batch():
    """Sub-batch sizes of the pipelined whole-job call (batch.ramp_sizes): positive, add up, small at both ends."""
    from audio_source_separation_b200.batch import ramp_sizes
    for n in list(range(1, 70)) + [100, 512, 1000]:
        sizes = ramp_sizes(n)
        assert sum(sizes) == n and all(v > 0 for v in sizes)
        if n >= 32:
            assert len(sizes) == 7 and sizes[0] <= sizes[3] and sizes[-1] <= sizes[3]
    assert ramp_sizes(64) == [4, 8, 12, 16, 12, 8, 4]


def test_bench_reference_arm_line_contract():
    """`bench.py --impl reference` (the CPU arm: the oracle port on the host cores) prints exactly one JSON line with the
    keys the driver reads; one timed update_once per worker keeps this to a few seconds."""
    import json
    env = dict(os.environ, RANK='0', WORLD_SIZE='1')
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--steps', '1', '--warmup', '0'],
                         capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    line = json.loads(lines[0])
    assert line['impl'] == 'reference' and line['unit'] == 'iterations/s' and line['higher_is_better'] is True
    assert line['value'] > 0 and line['cpu_baseline']['kind'] in ('reference', 'port') and line['cpu_baseline']['cores'] >= 1
    assert line['e2e'] == {'value': line['value'], 'unit': 'iterations/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}
    # a non-zero rank of a torchrun launch prints nothing and exits 0
    env['RANK'] = '1'
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--steps', '1', '--warmup', '0'],
                         capture_output=True, text=True, timeout=120, env=env, cwd=ROOT)
    assert out.returncode == 0 and out.stdout.strip() == ''


def test_batched_public_surface_is_complete():
    """The batch extension's public entry points (DESIGN.md section 6 / INTEGRATION.md): a refactoring once dropped some."""
    from audio_source_separation_b200.batch import BatchedGaussILRMA, gather_outputs, ramp_sizes, shard_range  # noqa: F401
    for name in ('open', 'reset', '__call__', 'separate_batch', 'separate_batch_sharded', 'separate_waveform_batch',
                 'separate_waveform_batch_sharded', 'separate_waveforms', 'update_once', 'compute_negative_loglikelihood',
                 'demix_filter', 'basis', 'activation'):
        assert hasattr(BatchedGaussILRMA, name), name


class _RecordingHandle:
    """Stand-in for `_lib.Handle` that records the order of the calls the pipelined driver makes (no device)."""
    log = []
    fail_on_size = None

    def __init__(self, **cfg):
        self.size = cfg['n_batch']
        self.priority = cfg['stream_priority']
        if self.size == _RecordingHandle.fail_on_size:
            raise RuntimeError("no device for a sub-batch of {}".format(self.size))

    def set_option(self, option, value):
        _RecordingHandle.log.append(('option', self.size, option, value))

    def reset_spatial(self):
        pass

    def set_state(self, which, value, dtype):
        import time
        time.sleep(0.002 * self.size)      # larger sub-batches are slower here, as on the host link: they would upload last

    def run(self, n):
        _RecordingHandle.log.append(('run', self.size, n))

    def launch_count(self):
        return 0

    def close(self):
        pass


def _pipelined_model(monkeypatch):
    from audio_source_separation_b200 import _lib, batch
    monkeypatch.setattr(_lib, 'Handle', _RecordingHandle)
    _RecordingHandle.log = []
    _RecordingHandle.fail_on_size = None
    model = batch.BatchedGaussILRMA.__new__(batch.BatchedGaussILRMA)
    model.n_basis, model.device, model.algorithm_spatial, model.normalize = 2, 0, 'IP', 'power'
    model.reference_id, model.domain, model.eps, model.threshold = 0, 2, 1e-12, 1e-12
    return model


def test_pipelined_job_uploads_inputs_in_sub_batch_order(monkeypatch):
    """`_pipelined` (profiles/round2_scaling.md): whatever the host threads' pace, the inputs go up in sub-batch order, every
    sub-batch asks for queued (not awaited) input copies, earlier sub-batches get the higher stream priority, and `on_done`
    is called in order on the calling thread."""
    import threading
    from audio_source_separation_b200 import _lib
    model = _pipelined_model(monkeypatch)
    sizes = [3, 1, 4, 2, 1]
    B, C, F, T = sum(sizes), 2, 5, 6
    fed, drained, done = [], [], []
    caller = threading.get_ident()

    def on_done(i, lo, hi):
        assert threading.get_ident() == caller
        done.append((i, lo, hi))

    model._pipelined(B, C, F, T, 7, np.zeros((B, C, F, 2)), np.zeros((B, C, 2, T)), sizes,
                     lambda h, lo, hi: fed.append((lo, hi)), lambda h, lo, hi: drained.append((lo, hi)), on_done=on_done)
    edges = np.concatenate(([0], np.cumsum(sizes)))
    spans = [(int(edges[i]), int(edges[i + 1])) for i in range(len(sizes))]
    assert fed == spans
    assert sorted(drained) == spans
    assert done == [(i,) + spans[i] for i in range(len(sizes))]
    assert [e for e in _RecordingHandle.log if e[0] == 'run'] and all(e[2] == 7 for e in _RecordingHandle.log if e[0] == 'run')
    asked = {(e[1], e[2]): e[3] for e in _RecordingHandle.log if e[0] == 'option'}
    assert all(asked[(n, _lib.OPT_ASYNC_INPUT)] == 1 and asked[(n, _lib.OPT_BLOCKING_SYNC)] == 1 for n in set(sizes))
    priorities = [slot[1].priority for slot in model._parts]
    assert priorities == sorted(priorities) and priorities[-1] == 0      # CUDA: lower number = higher priority
    assert [m['size'] for m in model.timeline] == sizes and all('out' in m for m in model.timeline)


def test_pipelined_job_does_not_hang_when_a_sub_batch_fails(monkeypatch):
    """A sub-batch that fails before its upload releases the next one's wait; the error reaches the caller."""
    model = _pipelined_model(monkeypatch)
    _RecordingHandle.fail_on_size = 4
    sizes = [2, 4, 3]
    B, C, F, T = sum(sizes), 2, 5, 6
    fed = []
    with pytest.raises(RuntimeError, match="no device for a sub-batch of 4"):
        model._pipelined(B, C, F, T, 3, np.zeros((B, C, F, 2)), np.zeros((B, C, 2, T)), sizes,
                         lambda h, lo, hi: fed.append((lo, hi)), lambda h, lo, hi: None)
    assert fed == [(0, 2), (6, 9)]


def test_every_entry_point_is_described_in_integration_md():
    """INTEGRATION.md section 3 names, for every symbol of include/bssgpu.h, the reference code it replaces (or says that
    there is none)."""
    doc = open(os.path.join(ROOT, 'INTEGRATION.md')).read()
    assert [s for s in _declared_symbols() if s not in doc] == []
