"""Phase surrogates on a long segment axis (multitaper pooling: 210 windows x 5 tapers = 1,050 terms): the
streamed-panel variant of the phase GEMM."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from multimodal_biosignal_analysis_b200 import kernels as K, synthetic as syn
from multimodal_biosignal_analysis_b200.signal_features import _dpss
dev = torch.device("cuda")
eeg, emg = syn.make_epochs(30, 8192, 64, 64)
starts = torch.from_numpy(syn.epoch_segment_starts(30, 8192, 2048, 1024)).to(dev)
tapers = torch.from_numpy(_dpss(2048, 3, 0.9).astype(np.float32)).to(dev)
X = K.fft_segments(torch.from_numpy(eeg).to(dev), starts, tapers, 0, 1, 100)
Y = K.fft_segments(torch.from_numpy(emg).to(dev), starts, tapers, 0, 1, 100)
L = X.shape[0] * X.shape[1]
res = K.csd_msc(X.view(L, 100, 64), Y.view(L, 100, 64))
n = 2000
for _ in range(2):
    ex, ms = K.surrogate_null(res, K.SURR_PHASE, 0, n, seed=3)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
ex, ms = K.surrogate_null(res, K.SURR_PHASE, 0, n, seed=3)
e1.record(); torch.cuda.synchronize()
t = e0.elapsed_time(e1)
kpb = (2 * L + 63) // 64 * 64
print(f"L = {L}: {n} phase surrogates in {t:.2f} ms = {n / t * 1e3 / 1e6:.3f} M surrogates/s, "
      f"{2.0 * n * 8192 * kpb * 100 / (t * 1e-3) / 1e12:.0f} TFLOP/s bf16 executed")
