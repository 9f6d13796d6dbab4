import sys, numpy as np, torch
sys.path.insert(0, '/root/repo')
from hallthrusterpem_b200.mc import HistogramSpec, MonteCarloMoments
from hallthrusterpem_b200.sampler import sample_inputs
from hallthrusterpem_b200.synthetic import spt100_batch
hist = HistogramSpec(angle_stride=8, sub_bits=3)
n, seed, A = 30000, 42, 200
def sampled(shift=(30.0, 0.8, 0.05)):
    a = MonteCarloMoments(n_angles=A, hist=hist, device=0, torr=133.322, scalar_shift=shift)
    a.accumulate_sampled(n, seed, 0); torch.cuda.synchronize(); return a.result().sums
def arrays(shift=(30.0, 0.8, 0.05)):
    b = MonteCarloMoments(n_angles=A, hist=hist, device=0, torr=133.322, scalar_shift=shift)
    b.accumulate(sample_inputs(n, seed, 0, device=0)); torch.cuda.synchronize(); return b.result().sums
def diff(x, y): return np.nonzero(x != y)[0]
s1, s2, s3 = sampled(), sampled(), sampled()
print('sampled vs sampled', diff(s1, s2), diff(s1, s3))
a1, a2 = arrays(), arrays()
print('arrays vs arrays', diff(a1, a2), 'sampled vs arrays', diff(s1, a1))
b0 = spt100_batch(20000, 4242 + 20000)
mc0 = MonteCarloMoments(n_angles=200, hist=hist, device=0, torr=133.322)
for lo, hi in ((0, 6666), (6666, 13333), (13333, 20000)):
    mc0.accumulate({k: torch.as_tensor(v[lo:hi], device='cuda:0') for k, v in b0.items()})
torch.cuda.synchronize()
s4, a4 = sampled(), arrays()
print('after mc0: sampled vs before', diff(s4, s1), 'arrays vs before', diff(a4, a1), 'sampled vs arrays', diff(s4, a4))
s5 = sampled(); print('again', diff(s5, s4))
