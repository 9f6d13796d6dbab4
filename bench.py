#!/usr/bin/env python
"""bench.py -- throughput of the fused cathode + plume hot path (sample x angle evaluations per second).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Headline workload (BASELINE.json configs[1]): SPT-100 plume+cathode Monte-Carlo, 1e6 samples x 200 angles PER GPU (weak
scaling: samples shard trivially, no data-path collective), fp64, sweep radius 1 m, all outputs materialised
(V_cc, j_ion (n, 200), div_angle, T_c).  A "step" is one pass of the hot path over that batch.

* `value`   : whole-job evals/s with inputs resident in HBM; every step is ONE hpem_eval() C-ABI call per rank, timed
              with CUDA events on the launching stream, max over ranks.  Footprint per step (1.74 GB) is far larger than
              L2 (126 MB), so no explicit flush is needed.
* `e2e`     : same metric through the public Python API with HOST buffers (pinned NumPy in, NumPy out): H2D of the 15
              inputs and D2H of every output inside the timed region.  `e2e_variants` adds what the reference's real
              caller sees: pageable inputs, the reference's own 91 angles, and the two outputs that never materialise
              j_ion (rank-6 SVD latents, probe log-likelihood).
* `roofline`: HBM-bound kernel; achieved = algorithmic bytes (8 + 144/A per eval) / measured kernel time, against
              MEASURED_PEAKS.json's copy bandwidth.
* `mc`      : the path the north star's multi-GPU sentence names -- reduce-only moments + histograms, inputs drawn on the
              device, ONE all-gather of the packed moments + fixed-order merge INSIDE the timed region.  Config 4 (1e8
              samples x 256 angles in total, strong scaling) and config 5 (1.25e8 samples x 512 angles per GPU, weak), each
              with a bit-for-bit check of counts and histograms against a single-GPU pass over the whole index range.
* `cfg3`    : BASELINE configs[2] at its stated size (1e7 samples x 256 angles per GPU, 20.5 GB of j_ion) with a strided
              spot check against the CPU oracle.
* `cpu_baseline` / `--impl reference`: the reference's NumPy path timed on the box's host cores -- the oracle port
              (bit-identical to the reference at A = 91) for the headline angle count, and the UNMODIFIED reference
              (oracle/_ref, staged by oracle/build_ref.py) at its hard-coded 91 angles.  The only uses of oracle/ here
              are those baselines and the cfg3 spot check.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

N_SAMPLES = 1_000_000
N_ANGLES = 200
METRIC = 'plume+cathode fp64 sample x angle evals/s'
UNIT = 'evals/s'
WORKLOAD = 'SPT-100 plume+cathode MC, 1e6 samples x 200 angles per GPU, fp64, r=1 m, j_ion materialised'
ALG_BYTES_PER_EVAL = 8.0 + 144.0 / N_ANGLES
K2_FP64_PER_EVAL = 10.0    # the reduce-only ALGORITHM per evaluation: 8 recurrence sweep incl. the two Simpson sums + sum + sum of squares
K2_FP64_EXECUTED = 8.6     # what the kernel executes per evaluation in its sweep since the Simpson sums come from a per-grid table
                           # (412 fp64 instructions per 16-angle chunk of a sample triple, ncu source page, profiles/r02_k2_ncu_summary.txt)
MC_SEED = 20240307


def workload_config(world: int) -> dict:
    """The `config` object of the JSON line -- identical, key for key, in both arms."""
    return {'workload': WORKLOAD, 'samples_per_gpu': N_SAMPLES, 'n_angles': N_ANGLES,
            'parallelism': f'samples sharded x{world}', 'sweep_radius_m': 1.0,
            'l2': 'per-step footprint 1.74 GB >> 126 MB L2, no flush needed'}


# ------------------------------------------------------------------------------------------------
# CPU side: the reference's NumPy path timed on host cores (cpu_baseline leg and --impl reference)
# ------------------------------------------------------------------------------------------------
def _cpu_chunk(args):
    seed, n, n_angles = args
    from hallthrusterpem_b200.synthetic import spt100_batch
    from oracle.ref_restated import cathode_coupling_oracle, current_density_oracle
    b = spt100_batch(n, seed)
    t0 = time.perf_counter()
    with np.errstate(all='ignore'):
        v = cathode_coupling_oracle(b)
        o = current_density_oracle(b, 1.0, n_angles)     # includes the reference's j_ion_coords loop (plume.py:152-155)
    dt = time.perf_counter() - t0
    return dt, float(o['j_ion'][0, 0]) + float(v['V_cc'][0])


def cpu_reference_rate(n_angles: int, target_seconds: float = 12.0, chunk: int = 2048):
    """Evals/s of the NumPy oracle with one worker process per available core (NumPy ufuncs are single-threaded).
    Bounded sample: every worker evaluates `tasks_per_worker` chunks of `chunk` samples."""
    import multiprocessing as mp
    cores = len(os.sched_getaffinity(0))
    ctx = mp.get_context('fork')
    with ctx.Pool(cores) as pool:
        pool.map(_cpu_chunk, [(i, 256, n_angles) for i in range(cores)])          # warm the workers
        t0 = time.perf_counter()
        pool.map(_cpu_chunk, [(100 + i, chunk, n_angles) for i in range(cores)])
        probe = time.perf_counter() - t0
        per_worker = max(1, int(target_seconds / max(probe, 1e-3)))
        tasks = [(1000 + i, chunk, n_angles) for i in range(cores * per_worker)]
        t0 = time.perf_counter()
        pool.map(_cpu_chunk, tasks, chunksize=1)
        wall = time.perf_counter() - t0
    n_samples = len(tasks) * chunk
    return {'value': n_samples * n_angles / wall, 'unit': UNIT, 'cores': cores, 'kind': 'port',
            'sample': f'{n_samples} samples x {n_angles} angles in chunks of {chunk} over {cores} worker processes '
                      f'({wall:.1f} s wall); NumPy restatement of plume.py:38-159 + cathode.py:24-38'}, wall, n_samples


def cpu_unmodified_reference_a91(target_seconds: float = 6.0):
    """The UNMODIFIED reference functions (hallmd.models.plume.current_density + cathode.cathode_coupling, hard-coded 91
    angles, plume.py:53) in ONE process / one thread -- how amisc calls them.  None when neither
    /root/reference nor the staged copy oracle/_ref exists."""
    try:
        from oracle import ref_import
        if not ref_import.available():
            return None
        current_density, cathode_coupling, _ = ref_import.load()
    except Exception as exc:  # noqa: BLE001
        return {'unavailable': f'{type(exc).__name__}: {exc}'}
    from hallthrusterpem_b200.synthetic import spt100_batch
    n = 100_000
    b = spt100_batch(n, 4)
    cat = {k: b[k] for k in ('P_b', 'V_a', 'T_e', 'V_vac', 'Pstar', 'P_T')}
    reps, t_used = 0, 0.0
    with np.errstate(all='ignore'):
        current_density(b)                                                       # warm
        while t_used < target_seconds and reps < 20:
            t0 = time.perf_counter()
            cathode_coupling(cat)
            current_density(b)
            t_used += time.perf_counter() - t0
            reps += 1
    return {'value': reps * n * 91 / t_used, 'unit': UNIT, 'cores': 1, 'kind': 'reference', 'n_angles': 91,
            'sample': f'{reps} x {n} samples x 91 angles, unmodified hallmd.models.plume.current_density + '
                      f'cathode.cathode_coupling in one process ({t_used:.1f} s)'}


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    steps_rates, wall_total = [], 0.0
    base = None
    for s in range(args.warmup + args.steps):
        base, wall, n_samples = cpu_reference_rate(N_ANGLES, target_seconds=6.0)
        if s >= args.warmup:
            steps_rates.append(base['value'])
            wall_total += wall
    value = float(np.mean(steps_rates))
    base['value'] = value
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': 1e3 * wall_total / max(1, args.steps), 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': workload_config(int(os.environ.get('WORLD_SIZE', '1'))),
        'note': 'each step is a bounded sample of the workload on the host cores (NumPy port of the reference at the '
                'headline 200 angles; the unmodified reference hard-codes 91 angles and is reported in reference_a91)',
        'cpu_baseline': base,
        'reference_a91': cpu_unmodified_reference_a91(),
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# clocks sampling during the timed region
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    QUERY = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
             'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
             'clocks_event_reasons.sw_power_cap')

    def __init__(self, index: int):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--id={index}', f'--query-gpu={self.QUERY}',
                                          '--format=csv,noheader,nounits', '-lms', '50'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def stop(self, t0: float, t1: float) -> dict:
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.1)
        self.proc.terminate()
        rows = [r for t, r in self.rows if t0 <= t <= t1]
        window = 'the timed steps and the ~0.6 s back-to-back loop of the same kernel that follows them'
        if not rows:
            rows, window = [r for _, r in self.rows], 'whole run (no sample fell inside the window)'
        sm, smax, reasons = [], None, set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for r in rows:
            f = [x.strip() for x in r.split(',')]
            try:
                sm.append(float(f[0]))
                smax = float(f[1])
                for nm, val in zip(names, f[3:7]):
                    if val.lower().startswith('active'):
                        reasons.add(nm)
            except (ValueError, IndexError):
                continue
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': smax, 'reasons': sorted(reasons),
                'samples': len(sm), 'window': window}


# ------------------------------------------------------------------------------------------------
# GPU side
# ------------------------------------------------------------------------------------------------
def _max_over_ranks(x: float, dev, world):
    import torch
    import torch.distributed as dist
    t = torch.tensor([x], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def bench_mc(name, n_total, n_angles, scaling, rank, world, dev, local_rank, barrier, fp64_peak, chunk=25_000_000):
    """Reduce-only Monte-Carlo (BASELINE configs 4 / 5): every rank accumulates its contiguous shard of ONE global sample index
    range (inputs drawn on the device), then ONE all-gather + fixed-order merge; both inside the CUDA-event region."""
    import torch
    from hallthrusterpem_b200 import _lib
    from hallthrusterpem_b200.mc import HistogramSpec, MonteCarloMoments
    from hallthrusterpem_b200.synthetic import shard_bounds
    lib = _lib.load()
    lo, hi = shard_bounds(n_total, world, rank)
    mc = MonteCarloMoments(n_angles=n_angles, hist=HistogramSpec(angle_stride=8), device=local_rank)
    mc.accumulate_sampled(min(chunk, hi - lo), MC_SEED, lo)        # warm-up: module load, workspace, pilot shifts, NCCL
    mc.merge()
    mc.reset()
    barrier()
    launches0 = lib.hpem_launch_count()
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    e0.record()
    for first in range(lo, hi, chunk):
        mc.accumulate_sampled(min(chunk, hi - first), MC_SEED, first)
    e1.record()
    mc.merge()
    e2.record()
    barrier()
    launches = lib.hpem_launch_count() - launches0
    ms_total = _max_over_ranks(e0.elapsed_time(e2), dev, world)
    ms_merge = _max_over_ranks(e1.elapsed_time(e2), dev, world)
    merged = mc.packed.clone()
    L = mc.layout
    out = None
    if rank == 0:
        # checker: ONE GPU over the whole index range; the sampler is index-addressed, so counts / histograms must be identical
        ref = MonteCarloMoments(n_angles=n_angles, hist=HistogramSpec(angle_stride=8), device=local_rank)
        for first in range(0, n_total, chunk):
            ref.accumulate_sampled(min(chunk, n_total - first), MC_SEED, first)
        torch.cuda.synchronize()
        a, b = merged.cpu().numpy(), ref.packed.cpu().numpy()
        counts_equal = bool(np.array_equal(a[:3], b[:3]) and np.array_equal(a[[3, 6, 9]], b[[3, 6, 9]]))
        hist_equal = bool(np.array_equal(a[L.off_hist:L.n_sums], b[L.off_hist:L.n_sums]))
        minmax_equal = bool(np.array_equal(a[L.n_sums:], b[L.n_sums:]))
        with np.errstate(all='ignore'):
            rel = np.abs(a[:L.off_hist] - b[:L.off_hist]) / np.maximum(np.abs(b[:L.off_hist]), 1e-300)
        res = mc.result()
        p = res.j_percentile([5, 50, 95])
        value = n_total * n_angles / (ms_total * 1e-3)
        out = {
            'config': name, 'workload': f'reduce-only MC, {n_total} samples x {n_angles} angles in total, inputs drawn on '
                                        f'device, histogram every 8th angle, {world} GPU(s)',
            'scaling': scaling, 'value': value, 'unit': UNIT, 'ms_total': ms_total, 'ms_allreduce': ms_merge,
            'allreduce_share': ms_merge / ms_total, 'collective': 'one all_gather of the packed [sums | minmax] vector + '
                                                                  'fixed-rank-order merge kernel (hpem_moments_merge)',
            'allreduce_bytes': int(L.n_packed * 8), 'nranks': world, 'gpu_launches_rank0': int(launches),
            'roofline': {'bound': 'fp64', 'alg_instr_per_eval': K2_FP64_PER_EVAL,
                         'achieved': K2_FP64_PER_EVAL * value / world / 1e12,
                         'peak': None if not fp64_peak else fp64_peak / 1e12, 'unit': 'T fp64 instr/s per GPU',
                         'frac': None if not fp64_peak else K2_FP64_PER_EVAL * value / world / fp64_peak,
                         'peak_source': 'measured DFMA issue peak (tools/fp64_peak.cu, profiles/fp64_peak.json)',
                         'alg_model': 'evals/s x 10 fp64 instructions (the reduce-only algorithm with both Simpson sums '
                                      'accumulated angle by angle: the definition of rounds 1-2, kept so that the fraction '
                                      'stays comparable) / measured DFMA issue peak',
                         'executed_sweep_instr_per_eval': K2_FP64_EXECUTED,
                         'frac_executed_sweep': None if not fp64_peak else K2_FP64_EXECUTED * value / world / fp64_peak,
                         'executed_note': 'the kernel now takes the Simpson sums from a per-grid table and executes 8.6 fp64 '
                                          'instructions per evaluation in its sweep (+ ~2.2 at 256 / ~1.2 at 512 angles for the '
                                          'per-sample part); ncu sm__pipe_fp64_cycles_active of the same kernel: '
                                          'profiles/r02_k2_ncu_summary.txt'},
            'check_vs_single_gpu': {'counts_equal': counts_equal, 'histograms_equal': hist_equal, 'minmax_equal': minmax_equal,
                                    'max_rel_diff_sums': float(np.nanmax(rel))},
            'stats': {'n_samples': res.n_samples, 'n_invalid': res.n_invalid, 'V_cc_mean': res.scalar('V_cc')['mean'],
                      'div_angle_mean': res.scalar('div_angle')['mean'],
                      'j_p5_p50_p95_at_0deg': [float(p[0, 0]), float(p[1, 0]), float(p[2, 0])]},
        }
    barrier()
    return out


def bench_cfg3(rank, world, dev, local_rank, barrier, peak):
    """BASELINE configs[2] at full size: 1e7 samples x 256 angles per GPU (20.5 GB of j_ion), device-resident, plus a strided
    spot check of the materialised output against the CPU oracle (checker only)."""
    import torch
    from hallthrusterpem_b200.engine import PreparedCall
    from hallthrusterpem_b200.synthetic import h9_sweep_batch
    n, A = 10_000_000, 256
    host = h9_sweep_batch(n, 20240307 + 3 + 1000 * rank)
    dev_in = {k: torch.as_tensor(v, device=dev) for k, v in host.items()}
    call = PreparedCall(dev_in, want_cathode=True, want_plume=True, sweep_radius=1.0, n_angles=A, extras=True)
    for _ in range(2):
        call.run()
    barrier()
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    evs[0].record()
    for i in range(3):
        call.run()
        evs[i + 1].record()
    barrier()
    ms = _max_over_ranks(evs[0].elapsed_time(evs[3]) / 3, dev, world)
    out = None
    if rank == 0:
        from oracle.ref_restated import cathode_coupling_oracle, current_density_oracle
        from tests import parity
        idx = np.arange(0, n, 4999)[:2048]
        sub = {k: v[idx] for k, v in host.items()}
        with np.errstate(all='ignore'):
            ref = current_density_oracle(sub, 1.0, A, 133.322, with_coords=False, return_internals=True)
            v_ref = cathode_coupling_oracle(sub, 133.322)['V_cc']
        call2 = PreparedCall(dev_in, want_cathode=True, want_plume=True, sweep_radius=1.0, n_angles=A, extras=True, torr=133.322)
        call2.run()
        torch.cuda.synchronize()
        tidx = torch.as_tensor(idx, device=dev)
        got = {k: v[tidx].cpu().numpy() for k, v in call2.result.items()}
        frac = parity.check_j_ion(got['j_ion'], ref['j_ion'], sub['I_B0'], [1.0], ref['_invalid'])
        parity.check_rel(got['cos_div'], ref['_cos_div'], 'cos_div')
        parity.check_rel(got['T_c'], ref['T_c'], 'T_c')
        parity.check_div_angle(got['div_angle'], ref['div_angle'], got['cos_div'], ref['_cos_div'])
        parity.check_rel(got['V_cc'], v_ref, 'V_cc', scale=parity.cathode_scale(sub, 133.322))
        assert np.array_equal(got['invalid'].astype(bool), ref['_invalid'])
        bpe = 8.0 + 144.0 / A
        achieved = bpe * n * A / (ms * 1e-3) / 1e9
        out = {'workload': f'H9-style pressure sweep incl. P_b = 0, {n} samples x {A} angles per GPU, all outputs materialised '
                           f'({n * A * 8 / 1e9:.1f} GB of j_ion per GPU)',
               'value': world * n * A / (ms * 1e-3), 'unit': UNIT, 'ms_per_step': ms,
               'roofline': {'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak,
                            'alg_bytes_per_eval': bpe},
               'oracle_spot_check': {'samples': int(idx.size), 'stride': 4999, 'rule': 'tests/parity.py (rel 1e-12)',
                                     'passed': True, 'frac_pure_rel_1e-12': frac}}
        del call2
    del call, dev_in
    torch.cuda.empty_cache()
    barrier()
    return out


def bench_e2e_variants(host, pinned, rank, world, dev, local_rank, barrier, cpu_extra):
    """End-to-end numbers through the public API with HOST buffers, beyond the headline `e2e`."""
    import torch
    from hallthrusterpem_b200.compression import SVD
    from hallthrusterpem_b200.likelihood import JionMeasurements, jion_log_likelihood
    from hallthrusterpem_b200.models import current_density, plume_cathode
    n = N_SAMPLES
    plume_keys = ('P_b', 'c0', 'c1', 'c2', 'c3', 'c4', 'c5', 'sigma_cex', 'I_B0')

    def timed(fn, steps=3):
        fn()
        barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            fn()
        torch.cuda.synchronize()
        return _max_over_ranks((time.perf_counter() - t0) / steps, dev, world)

    out = {}
    # (1) pageable inputs: what amisc actually passes (plain NumPy arrays)
    t = timed(lambda: float(plume_cathode(host, 1.0, n_angles=N_ANGLES, device=local_rank)['div_angle'][0]))
    out['pageable_inputs'] = {'value': world * n * N_ANGLES / t, 'unit': UNIT, 'n_angles': N_ANGLES, 'h2d_bytes_per_step': 8 * n * 15,
                              'd2h_bytes_per_step': 8 * n * (N_ANGLES + 3), 'api': 'plume_cathode(pageable NumPy dict) -> NumPy dict'}
    # (2) the reference's own angle count: current_density alone + cathode alone is what the YAML wires up; fused here
    t = timed(lambda: float(plume_cathode(host, 1.0, n_angles=91, device=local_rank)['div_angle'][0]))
    out['a91_pageable'] = {'value': world * n * 91 / t, 'unit': UNIT, 'n_angles': 91, 'h2d_bytes_per_step': 8 * n * 15,
                           'd2h_bytes_per_step': 8 * n * (91 + 3), 'api': 'plume_cathode(pageable NumPy dict, n_angles=91) -> NumPy dict',
                           'cpu_reference': cpu_extra.get('reference_a91')}
    # (3) rank-6 SVD latents of log10 j_ion (what amisc stores downstream, pem_v0_SPT-100.yml:272-280): 48 MB instead of 1.6 GB D2H
    plume_host = {k: host[k] for k in plume_keys}
    comp = SVD.from_samples({k: v[:500] for k, v in plume_host.items()}, n_angles=N_ANGLES, device=local_rank, rank=6)
    t = timed(lambda: float(comp.compress_inputs(plume_host)[0, 0]))
    out['latent_rank6'] = {'value': world * n * N_ANGLES / t, 'unit': UNIT, 'n_angles': N_ANGLES, 'h2d_bytes_per_step': 8 * n * 9,
                           'd2h_bytes_per_step': 8 * n * comp.rank, 'api': 'SVD.compress_inputs(NumPy dict) -> (n, 6) NumPy',
                           'cpu_port': cpu_extra.get('latent')}
    # (4) probe-angle Gaussian log-likelihood (mcmc.py:84-104): 8 MB D2H
    rng = np.random.default_rng(0)
    m = 64
    meas = JionMeasurements(rng.uniform(-1.5, 1.5, m), 10 ** rng.uniform(-2, 1, m), np.full(m, 0.1), n_angles=91, device=local_rank)
    t = timed(lambda: float(jion_log_likelihood(plume_host, meas)[0]))
    out['loglike_a91_m64'] = {'value': world * n * 91 / t, 'unit': UNIT, 'n_angles': 91, 'h2d_bytes_per_step': 8 * n * 9,
                              'd2h_bytes_per_step': 8 * n, 'api': 'jion_log_likelihood(NumPy dict, 64 probe points) -> (n,) NumPy',
                              'cpu_port': cpu_extra.get('loglike')}
    # (5) one process driving every visible GPU (device='all'): only meaningful at N = 1 launch with several devices visible
    if world == 1 and torch.cuda.device_count() > 1:
        ndev = torch.cuda.device_count()
        t = timed(lambda: float(current_density(host, 1.0, n_angles=N_ANGLES, device='all')['div_angle'][0]))
        out['single_process_all_gpus'] = {'value': n * N_ANGLES / t, 'unit': UNIT, 'devices': ndev,
                                          'api': "current_density(NumPy dict, device='all') -> NumPy dict"}
    return out


def cpu_extra_baselines():
    """Single-core NumPy timings of the two non-materialising consumers, on bounded samples (checker code, oracle/)."""
    from hallthrusterpem_b200.synthetic import spt100_batch
    from oracle.ref_restated import current_density_oracle
    out = {'reference_a91': cpu_unmodified_reference_a91()}
    n = 20_000
    b = spt100_batch(n, 9)
    rng = np.random.default_rng(0)
    U = np.linalg.qr(rng.normal(size=(N_ANGLES, 6)))[0]
    with np.errstate(all='ignore'):
        t0 = time.perf_counter()
        reps = 0
        while time.perf_counter() - t0 < 3.0:
            j = current_density_oracle(b, 1.0, N_ANGLES, with_coords=False)['j_ion']
            _ = np.log10(j) @ U
            reps += 1
        dt = time.perf_counter() - t0
    out['latent'] = {'value': reps * n * N_ANGLES / dt, 'unit': UNIT, 'cores': 1, 'kind': 'port',
                     'sample': f'{reps} x {n} samples x {N_ANGLES} angles: plume oracle + log10 + (A x 6) projection ({dt:.1f} s)'}
    try:
        from oracle.likelihood_oracle import jion_log_likelihood_oracle
        m = 64
        theta, y, sg = rng.uniform(-1.5, 1.5, m), 10 ** rng.uniform(-2, 1, m), np.full(m, 0.1)
        with np.errstate(all='ignore'):
            t0 = time.perf_counter()
            reps = 0
            while time.perf_counter() - t0 < 3.0:
                jion_log_likelihood_oracle(b, theta, y, sg, 91, 133.322)
                reps += 1
            dt = time.perf_counter() - t0
        out['loglike'] = {'value': reps * n * 91 / dt, 'unit': UNIT, 'cores': 1, 'kind': 'port',
                          'sample': f'{reps} x {n} samples x 91 angles, 64 probe points: plume oracle + interp1d + Gaussian sum ({dt:.1f} s)'}
    except Exception as exc:  # noqa: BLE001
        out['loglike'] = {'unavailable': f'{type(exc).__name__}: {exc}'}
    return out


def run_ours(args):
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))

    cpu_base, cpu_extra = None, {}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu_base, _, _ = cpu_reference_rate(N_ANGLES, target_seconds=12.0)     # before CUDA init (fork-safe)
        cpu_extra = cpu_extra_baselines()

    import torch
    import torch.distributed as dist
    from hallthrusterpem_b200 import _lib
    from hallthrusterpem_b200.engine import PreparedCall
    from hallthrusterpem_b200.synthetic import spt100_batch

    torch.cuda.set_device(local_rank)
    dev = f'cuda:{local_rank}'
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device(dev))
    lib = _lib.load()

    n, A = N_SAMPLES, N_ANGLES
    host = spt100_batch(n, 20240307 + 2 + 1000 * rank)            # each rank owns its own shard of the sample stream
    dev_in = {k: torch.as_tensor(v, device=dev) for k, v in host.items()}
    call = PreparedCall(dev_in, want_cathode=True, want_plume=True, sweep_radius=1.0, n_angles=A)
    stream = torch.cuda.current_stream()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank) if rank == 0 else None
    for _ in range(max(args.warmup, 3)):
        call.run()
    barrier()
    launches0 = lib.hpem_launch_count()
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    barrier()
    t_steps0 = time.perf_counter()
    evs[0].record(stream)
    for s in range(args.steps):
        call.run()
        evs[s + 1].record(stream)
    barrier()
    t_steps1 = time.perf_counter()
    launches = lib.hpem_launch_count() - launches0
    total_ms = evs[0].elapsed_time(evs[-1])
    step_ms = [evs[i].elapsed_time(evs[i + 1]) for i in range(args.steps)]
    total_ms_max = _max_over_ranks(total_ms, dev, world)
    value = world * n * A * args.steps / (total_ms_max * 1e-3)

    # The K timed steps last a few milliseconds -- shorter than one nvidia-smi sample.  Directly after them the SAME kernel
    # runs back to back for ~0.6 s, timed the same way: the clocks are sampled over steps + loop, and the loop gives the
    # sustained rate (on this part the 1000 W cap pulls the SM clock down within ~0.2 s of continuous fp64 + HBM work;
    # `value` is what one batch -- or a few -- gets, `sustained` what a long stream of batches gets).
    sus_n = max(200, int(0.6 / max(total_ms / args.steps * 1e-3, 1e-5)))
    es0, es1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    es0.record(stream)
    for _ in range(sus_n):
        call.run()
    es1.record(stream)
    barrier()
    t_sus1 = time.perf_counter()
    sus_ms = _max_over_ranks(es0.elapsed_time(es1), dev, world)
    clocks = sampler.stop(t_steps0, t_sus1) if sampler else None
    sustained = {'value': world * n * A * sus_n / (sus_ms * 1e-3), 'unit': UNIT, 'steps': sus_n, 'ms_per_step': sus_ms / sus_n,
                 'hbm_gbs': ALG_BYTES_PER_EVAL * n * A / (sus_ms / sus_n * 1e-3) / 1e9,
                 'note': 'same kernel, same buffers, back to back for ~0.6 s right after the timed steps'}

    # ---- e2e: public API, host buffers (pinned inputs), H2D + D2H inside the timed region, every rank ----
    from hallthrusterpem_b200.models import plume_cathode
    pinned = {k: torch.as_tensor(v).pin_memory().numpy() for k, v in host.items()}
    e2e_steps = max(2, min(args.steps, 5))
    out = plume_cathode(pinned, 1.0, n_angles=A, device=local_rank)            # warm: workspace + pinned pool
    del out
    out = plume_cathode(pinned, 1.0, n_angles=A, device=local_rank)
    h2d = 8 * n * 15
    d2h = sum(v.nbytes for k, v in out.items() if k != 'j_ion_coords')
    del out
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        out = plume_cathode(pinned, 1.0, n_angles=A, device=local_rank)
        checksum = float(out['div_angle'][0])                                  # result read on the host
        del out
    torch.cuda.synchronize()
    e2e_value = world * n * A * e2e_steps / _max_over_ranks(time.perf_counter() - t0, dev, world)
    e2e_variants = None if args.quick else bench_e2e_variants(host, pinned, rank, world, dev, local_rank, barrier, cpu_extra)
    del pinned

    peaks_path = ROOT / 'MEASURED_PEAKS.json'
    if peaks_path.exists():
        peak, peak_src = float(json.loads(peaks_path.read_text())['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
    else:
        peak, peak_src = 6650.0, 'fallback (B200_PROFILING.md)'
    fp64_peak = None
    fp_path = ROOT / 'profiles' / 'fp64_peak.json'
    if fp_path.exists():
        fp64_peak = json.loads(fp_path.read_text()).get('dfma_per_s_sustained')
    sustained['hbm_frac'] = sustained['hbm_gbs'] / peak

    # ---- the reduce-only Monte-Carlo path with its collective (BASELINE configs 4 and 5), and config 3 at full size ----
    mc_blocks, cfg3 = None, None
    if not args.quick:
        del call, dev_in
        torch.cuda.empty_cache()
        mc_blocks = [bench_mc('config 4', 100_000_000, 256, 'strong', rank, world, dev, local_rank, barrier, fp64_peak),
                     bench_mc('config 5', 125_000_000 * world, 512, 'weak', rank, world, dev, local_rank, barrier, fp64_peak)]
        cfg3 = bench_cfg3(rank, world, dev, local_rank, barrier, peak)

    if rank == 0:
        kernel_ms = float(np.mean(step_ms))                      # one kernel per step; events bracket each launch
        achieved = ALG_BYTES_PER_EVAL * n * A / (kernel_ms * 1e-3) / 1e9
        traffic = None
        tr_path = ROOT / 'profiles' / 'traffic.json'
        if tr_path.exists():
            traffic = json.loads(tr_path.read_text()).get('eval_uniform_kernel_1e6x200_bytes_per_launch')
        line = {
            'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': max(args.warmup, 3),
            'ms_per_step': total_ms_max / args.steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'f64', 'data': 'synthetic',
            'config': workload_config(world),
            'kernel': 'eval_uniform_kernel<plume,store> (one launch per step per rank)',
            'roofline': {'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak,
                         'traffic': traffic,
                         'traffic_source': 'profiles/traffic.json: dram__bytes_read.sum + dram__bytes_write.sum of one '
                                           '`ncu --set full` capture of this kernel at this size (a constant, not measured in this run)',
                         'peak_source': peak_src, 'alg_bytes_per_eval': ALG_BYTES_PER_EVAL, 'kernel_ms': kernel_ms,
                         'fp64_pipe': None if not fp64_peak else {
                             'executed_instr_per_eval': 11.0,
                             'frac_of_dfma_peak': 11.0 * n * A / (kernel_ms * 1e-3) / fp64_peak,
                             'note': 'non-binding axis: the recurrence kernel executes ~11 fp64-pipe instructions per evaluation '
                                     '(direct evaluation with two exp() per evaluation would need ~41)'}},
            'sustained': sustained,
            'cpu_baseline': cpu_base,
            'e2e': {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h,
                    'steps': e2e_steps, 'inputs': 'pinned',
                    'api': 'hallthrusterpem_b200.models.plume_cathode(NumPy dict) -> NumPy dict'},
            'e2e_variants': e2e_variants,
            'mc': mc_blocks,
            'cfg3': cfg3,
            'gpu_launches': int(launches),
            'clocks': clocks,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', choices=['ours', 'reference'], default='ours')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--quick', action='store_true', help='headline numbers only (skip e2e variants, mc and cfg3 blocks)')
    args = ap.parse_args()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()
