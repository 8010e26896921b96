import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from scipy import signal
from multimodal_biosignal_analysis_b200 import kernels as K, synthetic as syn
dev = torch.device("cuda")
eeg, emg = syn.make_epochs(30, 8192, 64, 64)
starts = torch.from_numpy(syn.epoch_segment_starts(30, 8192, 2048, 1024)).to(dev)
win = torch.from_numpy(signal.get_window("hann", 2048).astype(np.float32)[None]).to(dev)
X = K.fft_segments(torch.from_numpy(eeg).to(dev), starts, win, 1, 1, 100)[:, 0]
Y = K.fft_segments(torch.from_numpy(emg).to(dev), starts, win, 1, 1, 100)[:, 0]
res = K.csd_msc(X, Y)
sh = torch.from_numpy(np.random.default_rng(3).integers(1, len(starts), 1000).astype(np.int32)).to(dev)
for _ in range(2):
    ex, ms = K.surrogate_null(res, K.SURR_SHIFT, 0, 1000, shifts=sh)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    ex, ms = K.surrogate_null(res, K.SURR_SHIFT, 0, 1000, shifts=sh)
e1.record(); torch.cuda.synchronize()
print("shift null ms", e0.elapsed_time(e1) / 5, "checksum", int(ex.sum()), float(ms.double().sum()))
