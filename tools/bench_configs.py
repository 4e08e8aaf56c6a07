#!/usr/bin/env python
"""Secondary measurements: the other BASELINE.json configs (cfg1 EUC-NMF, cfg2 AuxIVA-IP, cfg3 single-mixture
Gauss-ILRMA, cfg4 FastMNMF) on one GPU, device time per update_once (CUDA events on the handle's stream, loop queued
without host synchronisation) next to the oracle port on the host for a few iterations.  One JSON line per config.
bench.py remains the headline measurement; this script only fills the table in profiles/.

    python tools/bench_configs.py [--skip-cpu] [--configs cfg1,cfg2,cfg3,cfg4,mnmf]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def device_loop_ms(handle, n_iter, warmup=3):
    handle.run(warmup)
    handle.synchronize()
    handle.timer_begin()
    handle.run(n_iter)
    ms = handle.timer_end()
    return ms / n_iter


def cpu_ms(step, n=2):
    step()
    t0 = time.perf_counter()
    for _ in range(n):
        step()
    return 1e3 * (time.perf_counter() - t0) / n


def cfg1(skip_cpu):
    from audio_source_separation_b200.algorithm.nmf import EUCNMF
    from oracle import nmf as o_nmf, synth
    Z = synth.spectrogram(257, 128, seed=0)
    np.random.seed(111)
    model = EUCNMF(n_basis=4)
    model.target = Z
    model._reset()
    model._prepare()
    ms = device_loop_ms(model._handle, 200, warmup=10)
    t0 = time.perf_counter()
    np.random.seed(111)
    EUCNMF(n_basis=4)(Z, iteration=50)
    call_ms = 1e3 * (time.perf_counter() - t0)
    out = {"config": "cfg1 EUC-NMF K=4 257x128", "gpu_ms_per_iter": ms, "gpu_it_per_s": 1e3 / ms,
           "gpu_call_50it_with_loss_ms": call_ms}
    if not skip_cpu:
        T, V = np.random.rand(257, 4), np.random.rand(4, 128)
        st = {'T': T, 'V': V}

        def step():
            st['T'], st['V'] = o_nmf.euc_step(Z, st['T'], st['V'])
        c = cpu_ms(step, 200)
        out.update(cpu_ms_per_iter=c, speedup=c / ms)
    return out


def cfg2(skip_cpu):
    from audio_source_separation_b200.bss.iva import AuxLaplaceIVA
    from oracle import auxiva as o_iva, synth
    X = synth.mix2(2, 1025, 256, seed=0)
    out = {"config": "cfg2 AuxLaplaceIVA 2ch 1025x256"}
    for spatial in ('IP', 'ISS', 'IP2'):
        model = AuxLaplaceIVA(algorithm_spatial=spatial, recordable_loss=False)
        model.input = X
        model._reset()
        ms = device_loop_ms(model._handle, 100, warmup=5)
        out["gpu_ms_per_iter_" + spatial] = ms
    ms = out["gpu_ms_per_iter_IP"]
    out["gpu_it_per_s"] = 1e3 / ms
    if not skip_cpu:
        st = o_iva.init_state(X)

        def step():
            o_iva.update_once(st, 'laplace', 'IP')
        c = cpu_ms(step, 5)
        out.update(cpu_ms_per_iter=c, speedup=c / ms)
    return out


def cfg3(skip_cpu):
    from audio_source_separation_b200.bss.ilrma import GaussILRMA, tILRMA
    from oracle import ilrma as o_ilrma, synth
    import warnings
    X = synth.mix2(4, 2049, 512, seed=0)
    W0, T0, V0 = synth.initial_state(4, 2049, 512, 2, seed=7)
    out = {"config": "cfg3 Gauss-ILRMA 4ch 2049x512 K=2, single mixture (L2 resident)"}
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        for name, kw in (('IP', {}), ('ISS', dict(algorithm_spatial='ISS')), ('IP2', dict(algorithm_spatial='IP2')),
                         ('IP_pb', dict(normalize='projection-back')), ('IP_K10', dict(n_basis=10))):
            K = kw.pop('n_basis', 2)
            model = GaussILRMA(n_basis=K, recordable_loss=False, **kw)
            model.input = X
            if K == 2:
                model._reset(demix_filter=W0, basis=T0, activation=V0)
            else:
                model._reset()
            out["gpu_ms_per_iter_" + name] = device_loop_ms(model._handle, 100, warmup=5)
        model = tILRMA(n_basis=2, nu=1000, recordable_loss=False)
        model.input = X
        model._reset(demix_filter=W0, basis=T0, activation=V0)
        out["gpu_ms_per_iter_tILRMA"] = device_loop_ms(model._handle, 100, warmup=5)
        # default user path: loss recorded every iteration, whole call
        t0 = time.perf_counter()
        GaussILRMA(n_basis=2)(X, iteration=100, demix_filter=W0, basis=T0, activation=V0)
        out["gpu_call_100it_with_loss_ms"] = 1e3 * (time.perf_counter() - t0)
    ms = out["gpu_ms_per_iter_IP"]
    out["gpu_it_per_s"] = 1e3 / ms
    out["cov_roofline_note"] = "3 x 33.6 MB per iteration from L2; see bench.py for the DRAM-bound batch"
    if not skip_cpu:
        st = o_ilrma.init_state(X, 2, W=W0, T=T0, V=V0)

        def step():
            o_ilrma.update_once(st)
        c = cpu_ms(step, 2)
        out.update(cpu_ms_per_iter=c, speedup=c / ms)
    return out


def cfg4(skip_cpu):
    from audio_source_separation_b200.bss.mnmf import FastMultichannelISNMF
    from oracle import fastmnmf as o_mnmf, synth
    M, F, T, K = 8, 2049, 1024, 2
    X = synth.mix2(M, F, T, seed=0)
    rng = np.random.default_rng(7)
    W0 = rng.random((M, F, K))
    H0 = rng.random((M, K, T))
    model = FastMultichannelISNMF(n_basis=K, recordable_loss=False)
    model.input = X
    model._reset(basis=W0, activation=H0)
    h = model._handle
    ms = device_loop_ms(h, 20, warmup=3)
    out = {"config": "cfg4 FastMNMF 8ch 2049x1024 K=2 N=8", "gpu_ms_per_iter": ms, "gpu_it_per_s": 1e3 / ms}
    loss = model.compute_negative_loglikelihood()
    out["loss_finite"] = bool(np.isfinite(loss))
    if not skip_cpu:
        Fs = 64   # the oracle needs (N,F,T,M) temporaries: time a 64-bin slice and scale by F / 64
        st = o_mnmf.init_state(X[:, :Fs], K, M, W=W0[:, :Fs], H=H0)

        def step():
            o_mnmf.update_once(st)
        c = cpu_ms(step, 1) * (F / Fs)
        out.update(cpu_ms_per_iter=c, cpu_note="oracle on a 64-bin slice x {:.1f}".format(F / Fs), speedup=c / ms)
    return out


def mnmf(skip_cpu):
    """Sawada IS-MNMF at the shape of the reference's MNMF notebook input: 2 channels,
    2049 bins, 256 frames, 2 sources, n_basis = 10 (the constructor default)."""
    from audio_source_separation_b200.bss.mnmf import MultichannelISNMF
    from oracle import mnmf as o_mnmf, synth
    C, N, F, T, K = 2, 2, 2049, 256, 10
    X = synth.mix2(C, F, T, seed=0)
    np.random.seed(111)
    model = MultichannelISNMF(n_basis=K, n_sources=N, recordable_loss=False)
    model.input = X
    model._reset()
    Z0, T0, V0 = model.latent.copy(), model.basis.copy(), model.activation.copy()
    ms = device_loop_ms(model._handle, 50, warmup=3)
    out = {"config": "IS-MNMF (Sawada) 2ch 2049x256 N=2 K=10", "gpu_ms_per_iter": ms, "gpu_it_per_s": 1e3 / ms}
    out["loss_finite"] = bool(np.isfinite(model.compute_negative_loglikelihood()))
    C4 = MultichannelISNMF(n_basis=K, n_sources=4, recordable_loss=False)
    C4.input = synth.mix2(4, F, 512, seed=0)
    C4._reset()
    out["gpu_ms_per_iter_4ch_512frames_N4"] = device_loop_ms(C4._handle, 20, warmup=3)
    if not skip_cpu:
        Fs = 128
        st = o_mnmf.init_state(X[:, :Fs], K, N, Z=Z0, T=T0[:Fs], V=V0)

        def step():
            o_mnmf.update_once(st)
        c = cpu_ms(step, 2) * (F / Fs)
        out.update(cpu_ms_per_iter=c, cpu_note="oracle on a 128-bin slice x {:.1f}".format(F / Fs), speedup=c / ms)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--skip-cpu', action='store_true')
    ap.add_argument('--configs', default='cfg1,cfg2,cfg3,cfg4')
    args = ap.parse_args()
    fns = {'cfg1': cfg1, 'cfg2': cfg2, 'cfg3': cfg3, 'cfg4': cfg4, 'mnmf': mnmf}
    for name in args.configs.split(','):
        print(json.dumps(fns[name](args.skip_cpu)), flush=True)


if __name__ == '__main__':
    main()
