#!/usr/bin/env python
"""Per-row measurement of the SURVEY section-8 rows beyond the bench.py headline (one JSON line per row).

    python tools/bench_rows.py [--no-cpu] > profiles/rNN_rows_bench.jsonl

Every line: {"row", "kernel", "workload", "value", "unit", "ms", "roofline": {bound, achieved, peak, unit, frac, alg},
"cpu_baseline": {value, unit, cores, kind, sample}}.  GPU times are CUDA-event medians on torch's current stream (the
stream every launch below is issued on), inputs resident in HBM.  The CPU leg times the oracle (NumPy restatement of the
reference, 1 process) on a bounded sample of the same workload -- the only use of oracle/ here.
Roofline denominators: MEASURED_PEAKS.json hbm_gbs (else 6650 GB/s fallback) and profiles/fp64_peak.json (measured DFMA
issue rate).  `alg` states the algorithmic bytes / fp64 instructions per unit used for `achieved`.
"""
from __future__ import annotations

import argparse
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
TORR = 133.322


def peaks():
    p = ROOT / 'MEASURED_PEAKS.json'
    hbm = float(json.loads(p.read_text())['hbm_gbs']) if p.exists() else 6650.0
    f = ROOT / 'profiles' / 'fp64_peak.json'
    fp64 = float(json.loads(f.read_text())['dfma_per_s_sustained']) if f.exists() else 1.85e13
    return hbm, fp64


def gpu_ms(fn, reps=7, warm=3):
    import torch
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for _ in range(reps):
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


def cpu_rate(fn, units, target_s=4.0):
    """units per second of `fn` (one call = `units` units), repeated for ~target_s."""
    fn()
    t0 = time.perf_counter()
    reps = 0
    while time.perf_counter() - t0 < target_s:
        fn()
        reps += 1
    wall = time.perf_counter() - t0
    return units * reps / wall, reps, wall


def emit(row, kernel, workload, units, unit, ms, bound, alg_per_unit, alg_text, cpu=None):
    hbm, fp64 = peaks()
    if bound == 'hbm':
        achieved, peak, u = alg_per_unit * units / (ms * 1e-3) / 1e9, hbm, 'GB/s'
    else:
        achieved, peak, u = alg_per_unit * units / (ms * 1e-3) / 1e12, fp64 / 1e12, 'T fp64 instr/s'
    print(json.dumps({'row': row, 'kernel': kernel, 'workload': workload, 'value': units / (ms * 1e-3), 'unit': unit, 'ms': ms,
                      'roofline': {'bound': bound, 'achieved': achieved, 'peak': peak, 'unit': u, 'frac': achieved / peak,
                                   'alg': alg_text}, 'cpu_baseline': cpu}), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--no-cpu', action='store_true')
    args = ap.parse_args()
    import torch
    from hallthrusterpem_b200.compression import SVD
    from hallthrusterpem_b200.engine import PreparedCall
    from hallthrusterpem_b200.likelihood import JionMeasurements, jion_log_likelihood
    from hallthrusterpem_b200.mc import HistogramSpec, MonteCarloMoments
    from hallthrusterpem_b200.sampler import sample_inputs
    from hallthrusterpem_b200.synthetic import h9_sweep_batch, spt100_batch
    from oracle.compression_oracle import compress_oracle, normalize_log10
    from oracle.likelihood_oracle import jion_log_likelihood_oracle
    from oracle.ref_restated import cathode_coupling_oracle, current_density_oracle

    def dev(b):
        return {k: torch.as_tensor(v, device='cuda:0') for k, v in b.items()}

    def cpu(fn, units, unit, sample):
        if args.no_cpu:
            return None
        with np.errstate(all='ignore'):
            v, reps, wall = cpu_rate(fn, units)
        return {'value': v, 'unit': unit, 'cores': 1, 'kind': 'port', 'sample': f'{sample}, {reps} repetitions in {wall:.1f} s'}

    # ---- (a) materialising path at the other BASELINE angle counts (parity cases of bench.py's workload) ----
    for label, n, A, gen in (('a/cfg1', 1000, 100, spt100_batch), ('a/ref-A91', 1_000_000, 91, spt100_batch),
                             ('a/cfg3-H9', 4_000_000, 256, h9_sweep_batch), ('a/cfg5-A512', 1_000_000, 512, spt100_batch)):
        b = gen(n, 7)
        call = PreparedCall(dev(b), want_cathode=True, want_plume=True, sweep_radius=1.0, n_angles=A)
        ms = gpu_ms(call.run)
        small = {k: v[:20000] for k, v in b.items()}
        c = cpu(lambda: (cathode_coupling_oracle(small, TORR), current_density_oracle(small, 1.0, A, TORR)), min(n, 20000) * A,
                'evals/s', f'{min(n, 20000)} samples x {A} angles')
        emit(label, 'eval_uniform_kernel', f'{n} samples x {A} angles, all outputs materialised', n * A,
             'evals/s', ms, 'hbm', 8 + 144 / A, f'{8 + 144 / A:.3f} B/eval (8 + 144/A)', c)

    # ---- (f1) multi-radius sweep ----
    n, A, R = 400_000, 91, 25      # 7.3 GB of j_ion: several waves of blocks (1e5 samples is under two)
    b = spt100_batch(n, 8, c3_test_range=True)
    radii = np.linspace(1.0, 1.2, R)
    call = PreparedCall(dev(b), want_cathode=False, want_plume=True, sweep_radius=radii, n_angles=A)
    ms = gpu_ms(call.run)
    small = {k: v[:2000] for k, v in b.items()}
    c = cpu(lambda: current_density_oracle(small, radii, A, TORR), 2000 * A * R, 'sample x angle x radius /s', f'2000 samples x {A} x {R}')
    per = 8 + (144 + 16 * R) / (A * R)
    emit('f1', 'eval_radii_stream_kernel', f'{n} samples x {A} angles x {R} radii (tests/test_plume.py:31-35 shape)', n * A * R,
         'sample x angle x radius /s', ms, 'hbm', per, f'{per:.3f} B per (sample, angle, radius)', c)

    # ---- (f2) interpolation to probe angles + Gaussian log-likelihood ----
    n, A, m = 1_000_000, 91, 64
    b = spt100_batch(n, 9)
    rng = np.random.default_rng(0)
    theta, y = rng.uniform(-1.5, 1.5, m), 10 ** rng.uniform(-2, 1, m)
    sig = 0.2 * y + 0.01
    meas = JionMeasurements(theta, y, sig, n_angles=A, device=0)
    d = dev({k: v for k, v in b.items() if k in ('P_b', 'c0', 'c1', 'c2', 'c3', 'c4', 'c5', 'sigma_cex', 'I_B0')})
    ms = gpu_ms(lambda: jion_log_likelihood(d, meas, torr=TORR))
    small = {k: v[:20000] for k, v in b.items()}
    c = cpu(lambda: jion_log_likelihood_oracle(small, theta, y, sig, A, TORR), 20000 * A, 'evals/s', f'20000 samples x {A} angles, {m} probe points')
    emit('f2', 'loglike_kernel', f'{n} samples x {A} angles, {m} probe points, log-likelihood only', n * A, 'evals/s', ms, 'fp64',
         8 + 5 * m / A, f'{8 + 5 * m / A:.2f} fp64 instr/eval (8 recurrence sweep + 5 per probe point / A)', c)

    # ---- (f3) on-device prior sampler, and the sampled reduce-only pass ----
    n = 8_000_000
    ms = gpu_ms(lambda: sample_inputs(n, 11, 0, device=0), reps=5)
    g = np.random.default_rng(1)
    c = cpu(lambda: [g.uniform(0, 1, 200_000) for _ in range(12)] + [np.exp(g.uniform(0, 1, 200_000)) for _ in range(3)], 200_000,
            'samples/s', '200000 samples x 15 inputs, numpy Generator.uniform (+exp for the 3 LogUniform inputs)')
    emit('f3', 'sample_inputs_kernel', f'{n} samples x 15 inputs (SPT-100 priors)', n, 'samples/s', ms, 'hbm', 120.0, '120 B/sample (15 fp64 stores)', c)
    for A, stride in ((256, 8), (512, 8), (91, 8)):
        n = 40_000_000
        mc = MonteCarloMoments(n_angles=A, device=0, hist=HistogramSpec(angle_stride=stride))
        ms = gpu_ms(lambda: mc.accumulate_sampled(n, 7, 0), reps=5, warm=2)
        emit('f3+e/K2', 'moments_kernel<sampled>', f'reduce-only MC, {n} samples x {A} angles, inputs drawn on device, histogram every '
             f'{stride}th angle', n * A, 'evals/s', ms, 'fp64', 10.0, '10 fp64 instr/eval: the reduce-only algorithm (8 recurrence sweep incl. the two Simpson sums + sum + sum of squares); '
             'the kernel executes 8.8 in its sweep since the Simpson sums come from the per-grid table')

    # ---- (f4) SVD compression ----
    for A in (91, 200):
        n = 1_000_000
        b = {k: v for k, v in spt100_batch(n, 12).items() if k != 'T'}
        fit = {k: v[:500] for k, v in b.items()}
        comp = SVD.from_samples(fit, n_angles=A, torr=TORR, device=0, reconstruction_tol=0.01)
        d = dev(b)
        ms = gpu_ms(lambda: comp.compress_inputs(d, torr=TORR))
        small = {k: v[:20000] for k, v in b.items()}
        c = cpu(lambda: compress_oracle(comp.projection_matrix, normalize_log10(current_density_oracle(small, 1.0, A, TORR, with_coords=False)['j_ion'])),
                20000 * A, 'evals/s', f'20000 samples x {A} angles, plume oracle + log10 + projection (rank {comp.rank})')
        emit('f4', 'latent_kernel', f'{n} samples x {A} angles -> rank {comp.rank} latent (fused plume + log10 + projection)', n * A, 'evals/s', ms,
             'fp64', 8 + 30 + comp.rank, f'{8 + 30 + comp.rank} fp64 instr/eval (8 sweep + ~30 log10 + rank FMAs)', c)
        z = comp.compress_inputs(d, torr=TORR)
        ms = gpu_ms(lambda: comp.reconstruct_field(z))
        zs = z[:20000].cpu().numpy()
        c = cpu(lambda: 10.0 ** (zs @ comp.projection_matrix.T), 20000 * A, 'elements/s', f'20000 x {A}, NumPy matmul + 10**x')
        emit('f4', 'reconstruct_kernel', f'{n} x rank {comp.rank} latent -> {n} x {A} field', n * A, 'elements/s', ms, 'hbm', 8 + 8 * comp.rank / A,
             f'{8 + 8 * comp.rank / A:.2f} B/element (one fp64 store + latent reads)', c)


if __name__ == '__main__':
    main()
