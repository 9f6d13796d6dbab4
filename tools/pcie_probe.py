#!/usr/bin/env python
"""Developer probe: host<->device copy bandwidth per rank, alone and concurrently (explains the e2e scaling of bench.py).
   python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29611 tools/pcie_probe.py"""
import os
import subprocess
import time

import torch
import torch.distributed as dist

rank, world, local = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1)), int(os.environ.get('LOCAL_RANK', 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group('nccl', device_id=torch.device(f'cuda:{local}'))
nbytes = 1600 * 1000 * 1000
d = torch.empty(nbytes, dtype=torch.uint8, device=f'cuda:{local}')
h = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)


def bar():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def d2h(reps=5):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        h.copy_(d, non_blocking=True)
    torch.cuda.synchronize()
    return nbytes * reps / (time.perf_counter() - t0) / 1e9


def h2d(reps=5):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    return nbytes * reps / (time.perf_counter() - t0) / 1e9


d2h(1); h2d(1)
if rank == 0:
    print(subprocess.run(['nvidia-smi', 'topo', '-m'], capture_output=True, text=True).stdout, flush=True)
    print('cpus', len(os.sched_getaffinity(0)), sorted(os.sched_getaffinity(0))[:4], '...', flush=True)
    try:
        print('numa nodes', sorted(os.listdir('/sys/devices/system/node'))[:8], flush=True)
    except Exception as exc:  # noqa: BLE001
        print('numa: ', exc)
for r in range(world):          # one rank at a time
    bar()
    if r == rank:
        print(f'rank {rank} alone: D2H {d2h():.1f} GB/s  H2D {h2d():.1f} GB/s', flush=True)
bar()
a, b = d2h(), None
bar()
b = h2d()
t = torch.tensor([a, b], dtype=torch.float64, device=f'cuda:{local}')
if world > 1:
    lst = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(lst, t)
else:
    lst = [t]
if rank == 0:
    print('concurrent D2H per rank:', [round(float(x[0]), 1) for x in lst], 'sum', round(sum(float(x[0]) for x in lst), 1), flush=True)
    print('concurrent H2D per rank:', [round(float(x[1]), 1) for x in lst], 'sum', round(sum(float(x[1]) for x in lst), 1), flush=True)
if world > 1:
    dist.destroy_process_group()
