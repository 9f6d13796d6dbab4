#!/usr/bin/env python
"""bench.py -- throughput of the fused cathode + plume hot path (sample x angle evaluations per second).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[1]): SPT-100 plume+cathode Monte-Carlo, 1e6 samples x 200 angles PER GPU (weak
scaling: samples shard trivially, no data-path collective), fp64, sweep radius 1 m, all outputs materialised
(V_cc, j_ion (n, 200), div_angle, T_c).  A "step" is one pass of the hot path over that batch.

* `value`  : whole-job evals/s with inputs resident in HBM; every step is ONE hpem_eval() C-ABI call per rank,
             timed with CUDA events on the launching stream, max over ranks.  Footprint per step (1.74 GB) is far
             larger than L2 (126 MB), so no explicit flush is needed.
* `e2e`    : same metric through the public Python API with HOST buffers (pinned NumPy in, NumPy out): H2D of the
             15 inputs and D2H of every output inside the timed region.
* `roofline`: HBM-bound kernel; achieved = algorithmic bytes (8 + 144/A per eval) / measured kernel time, against
             MEASURED_PEAKS.json's copy bandwidth.  `fp64` adds the second (non-binding) roofline.
* `cpu_baseline` / `--impl reference`: the oracle (NumPy restatement of the reference, bit-identical to it at A=91)
             timed on the box's host cores -- the ONLY use of oracle/ in this file.
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
ALG_FP64_PER_EVAL = 41.0                             # BASELINE.md section 4 (direct evaluation)


# ------------------------------------------------------------------------------------------------
# CPU side: the oracle timed on host cores (cpu_baseline leg and --impl reference)
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
        'config': {'workload': WORKLOAD, 'n_angles': N_ANGLES, 'note': 'each step is a bounded sample of the workload'},
        'cpu_baseline': base,
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
                                          '--format=csv,noheader,nounits', '-lms', '100'],
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
        time.sleep(0.15)
        self.proc.terminate()
        rows = [r for t, r in self.rows if t0 - 0.05 <= t <= t1 + 0.15] or [r for _, r in self.rows]
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
                'samples': len(sm)}


# ------------------------------------------------------------------------------------------------
# GPU side
# ------------------------------------------------------------------------------------------------
def run_ours(args):
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))

    cpu_base = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu_base, _, _ = cpu_reference_rate(N_ANGLES, target_seconds=12.0)     # before CUDA init (fork-safe)

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
    if sampler:
        time.sleep(0.3)
    t_load0 = time.perf_counter()
    for _ in range(max(args.warmup, 3)):
        call.run()
    barrier()
    launches0 = lib.hpem_launch_count()
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    barrier()
    t_wall0 = time.perf_counter()
    evs[0].record(stream)
    for s in range(args.steps):
        call.run()
        evs[s + 1].record(stream)
    barrier()
    t_wall1 = time.perf_counter()
    launches = lib.hpem_launch_count() - launches0
    total_ms = evs[0].elapsed_time(evs[-1])
    step_ms = [evs[i].elapsed_time(evs[i + 1]) for i in range(args.steps)]
    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms_max = float(t.item())
    value = world * n * A * args.steps / (total_ms_max * 1e-3)

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
    t_e2e = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
    e2e_value = world * n * A * e2e_steps / float(t_e2e.item())

    # the device-timed region lasts only milliseconds; clocks are sampled (100 ms period) from the first warm-up launch to
    # the end of the e2e loop, all of which keeps the GPU busy with the same kernels
    clocks = sampler.stop(t_load0, time.perf_counter()) if sampler else None

    if rank == 0:
        peaks_path = ROOT / 'MEASURED_PEAKS.json'
        if peaks_path.exists():
            peak, peak_src = float(json.loads(peaks_path.read_text())['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
        else:
            peak, peak_src = 6650.0, 'fallback (B200_PROFILING.md)'
        kernel_ms = float(np.mean(step_ms))                      # one kernel per step; events bracket each launch
        achieved = ALG_BYTES_PER_EVAL * n * A / (kernel_ms * 1e-3) / 1e9
        traffic = None
        tr_path = ROOT / 'profiles' / 'traffic.json'
        if tr_path.exists():
            traffic = json.loads(tr_path.read_text()).get('eval_uniform_kernel_1e6x200_bytes_per_launch')
        fp64_peak = None
        fp_path = ROOT / 'profiles' / 'fp64_peak.json'
        if fp_path.exists():
            fp64_peak = json.loads(fp_path.read_text()).get('dfma_per_s_sustained')
        line = {
            'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': max(args.warmup, 3),
            'ms_per_step': total_ms_max / args.steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'f64', 'data': 'synthetic',
            'config': {'workload': WORKLOAD, 'samples_per_gpu': n, 'n_angles': A, 'parallelism': f'samples sharded x{world}',
                       'l2': 'per-step footprint 1.74 GB >> 126 MB L2, no flush needed',
                       'kernel': 'eval_uniform_kernel<plume,store> (one launch per step per rank)'},
            'roofline': {'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak,
                         'traffic': traffic, 'peak_source': peak_src, 'alg_bytes_per_eval': ALG_BYTES_PER_EVAL,
                         'kernel_ms': kernel_ms,
                         'fp64': None if not fp64_peak else {
                             'alg_instr_per_eval': ALG_FP64_PER_EVAL,
                             'achieved_instr_per_s': ALG_FP64_PER_EVAL * n * A / (kernel_ms * 1e-3),
                             'peak_instr_per_s': fp64_peak,
                             'frac': ALG_FP64_PER_EVAL * n * A / (kernel_ms * 1e-3) / fp64_peak,
                             'note': 'non-binding; measured DFMA issue peak (tools/fp64_peak.cu). frac > 1 because the Gaussian '
                                     'recurrence executes ~11 fp64 instr/eval where direct evaluation (2 exp/eval, the '
                                     'algorithmic figure of BASELINE.md) needs 41'}},
            'cpu_baseline': cpu_base,
            'e2e': {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h,
                    'steps': e2e_steps, 'api': 'hallthrusterpem_b200.models.plume_cathode(NumPy dict) -> NumPy dict'},
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
    args = ap.parse_args()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()
