#!/usr/bin/env python
"""ncu target: one K2 launch (arrays, hist stride from argv, A from argv)."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from hallthrusterpem_b200.mc import HistogramSpec, MonteCarloMoments
from hallthrusterpem_b200.synthetic import spt100_batch
A = int(sys.argv[1]) if len(sys.argv) > 1 else 256
stride = int(sys.argv[2]) if len(sys.argv) > 2 else 8
n = int(float(sys.argv[3])) if len(sys.argv) > 3 else 2_000_000
b = {k: torch.as_tensor(v, device='cuda:0') for k, v in spt100_batch(n, 1).items()}
mc = MonteCarloMoments(n_angles=A, device=0, hist=HistogramSpec(angle_stride=stride))
mc.accumulate(b)
mc.accumulate(b)
torch.cuda.synchronize()
print('done')
