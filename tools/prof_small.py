import sys
from pathlib import Path; sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from hallthrusterpem_b200.engine import PreparedCall
from hallthrusterpem_b200.synthetic import spt100_batch
b = {k: torch.as_tensor(v, device='cuda:0') for k, v in spt100_batch(1_000_000, 1).items()}
call = PreparedCall(b, want_cathode=True, want_plume=True, sweep_radius=1.0, n_angles=64, want_j_ion=False)
for _ in range(3): call.run()
torch.cuda.synchronize()
