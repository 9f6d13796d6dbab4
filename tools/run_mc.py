#!/usr/bin/env python
"""Reduce-only Monte-Carlo run for BASELINE.json configs 4-5 (1e8-1e9 samples, no j_ion materialisation).

    python tools/run_mc.py --samples 1e8 --angles 256                      # one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29700 \
        tools/run_mc.py --samples 1e9 --angles 512 [--weak]                # samples sharded over the ranks

Inputs are drawn on the fly by the on-device sampler (global sample index space, shard-invariant), each rank
accumulates its contiguous index range in chunks, and ONE all-gather of the packed moments + a fixed-rank-order merge
kernel combine the ranks (NCCL).  Rank 0 prints a JSON line with throughput (device-timed, max over ranks) and a few statistics.
"""
import argparse
import json
import os
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--samples', type=float, default=1e8, help='total samples (per rank with --weak)')
    ap.add_argument('--angles', type=int, default=256)
    ap.add_argument('--chunk', type=float, default=4e6, help='samples per kernel launch')
    ap.add_argument('--seed', type=int, default=20240307)
    ap.add_argument('--weak', action='store_true')
    ap.add_argument('--hist-stride', type=int, default=8)
    args = ap.parse_args()

    import torch
    import torch.distributed as dist
    from hallthrusterpem_b200 import _lib
    from hallthrusterpem_b200.mc import HistogramSpec, MonteCarloMoments
    from hallthrusterpem_b200.synthetic import shard_bounds

    rank, world = int(os.environ.get('RANK', '0')), int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device(f'cuda:{local}'))
    n_total = int(args.samples) * (world if args.weak else 1)
    lo, hi = shard_bounds(n_total, world, rank)
    mc = MonteCarloMoments(n_angles=args.angles, hist=HistogramSpec(angle_stride=args.hist_stride), device=local)
    chunk = int(args.chunk)
    mc.accumulate_sampled(min(chunk, hi - lo), args.seed, lo)      # warm-up (workspace, module load), then reset
    mc.reset()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    launches0 = _lib.load().hpem_launch_count()
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    t0 = time.perf_counter()
    e0.record()
    for first in range(lo, hi, chunk):
        mc.accumulate_sampled(min(chunk, hi - first), args.seed, first)
    e1.record()
    mc.merge()
    e2.record()
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    t = torch.tensor([e0.elapsed_time(e2), e1.elapsed_time(e2)], dtype=torch.float64, device=f'cuda:{local}')
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        res = mc.result()
        total_ms, merge_ms = float(t[0]), float(t[1])
        p = res.j_percentile([5, 50, 95]) if args.hist_stride > 0 else np.full((3, 1), np.nan)
        print(json.dumps({
            'workload': f'reduce-only MC, {n_total} samples x {args.angles} angles, sampled on device, {world} GPU(s)',
            'value': n_total * args.angles / (total_ms * 1e-3), 'unit': 'evals/s', 'n_gpus': world, 'ms_total': total_ms,
            'ms_allreduce': merge_ms, 'wall_s': wall, 'launches_rank0': int(_lib.load().hpem_launch_count() - launches0),
            'allreduce_bytes': int(mc.packed.numel() * 8), 'n_samples': res.n_samples, 'n_invalid': res.n_invalid,
            'V_cc': res.scalar('V_cc'), 'div_angle': res.scalar('div_angle'), 'T_c': res.scalar('T_c'),
            'j_mean_0_mid_end': [float(res.j_mean[0]), float(res.j_mean[args.angles // 2]), float(res.j_mean[-1])],
            'j_p5_p50_p95_at_angle0': [float(p[0, 0]), float(p[1, 0]), float(p[2, 0])],
        }), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
