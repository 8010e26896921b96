"""One config-2 launch pair of the tensor-core Welch kernel for ncu (second launch is the profiled one)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from multimodal_biosignal_analysis_b200 import kernels as K, synthetic as syn

dev = torch.device("cuda:0")
eeg, emg = syn.make_epochs(30, 8192, 64, 64, seed=20260102)
st = syn.epoch_segment_starts(30, 8192, 2048, 1024)
e_d, m_d = torch.from_numpy(eeg).to(dev), torch.from_numpy(emg).to(dev)
plan = K.WelchHannPlan(st, 2048, 1, 100)
spec = torch.empty((len(st), 1, 100, 128), dtype=torch.complex64, device=dev)
for _ in range(3):
    plan.spectra(e_d, spec[..., :64], m_d, spec[..., 64:])
torch.cuda.synchronize()
print("done")
