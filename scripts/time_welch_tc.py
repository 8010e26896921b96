"""Timing of the tensor-core Welch kernel on config 2 (CUDA events, inputs alternate between two resident recordings)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from multimodal_biosignal_analysis_b200 import kernels as K, synthetic as syn

dev = torch.device("cuda:0")
NE = int(os.environ.get("NE", "30"))
eeg, emg = syn.make_epochs(NE, 8192, 64, 64, seed=20260102)
st = syn.epoch_segment_starts(NE, 8192, 2048, 1024)
sets = [(torch.from_numpy(eeg).to(dev), torch.from_numpy(emg).to(dev)),
        (torch.from_numpy(eeg[::-1].copy()).to(dev), torch.from_numpy(emg[::-1].copy()).to(dev))]
plan = K.WelchHannPlan(st, 2048, 1, 100)
spec = torch.empty((len(st), 1, 100, 128), dtype=torch.complex64, device=dev)
for _ in range(5):
    plan.spectra(sets[0][0], spec[..., :64], sets[0][1], spec[..., 64:])
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):                       # 10 launches per replay: no host work between kernels
    for i in range(10):
        a, b = sets[i & 1]
        plan.spectra(a, spec[..., :64], b, spec[..., 64:])
g.replay()
torch.cuda.synchronize()
e0.record()
for i in range(5):
    g.replay()
e1.record()
torch.cuda.synchronize()
print(f"epochs {NE} half blocks {plan.n_half_blocks} CMC_DT_DBG={os.environ.get('CMC_DT_DBG', '0')}: {e0.elapsed_time(e1) / 50 * 1e3:.1f} us per launch")
