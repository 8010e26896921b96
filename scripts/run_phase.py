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
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2560
for _ in range(2):
    ex, ms = K.surrogate_null(res, K.SURR_PHASE, 0, n, seed=7)
torch.cuda.synchronize()
print("ok", float(ms.max()))
