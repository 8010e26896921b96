import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from multimodal_biosignal_analysis_b200 import kernels as K, synthetic as syn
from multimodal_biosignal_analysis_b200.signal_features import _dpss
dev = torch.device("cuda")
eeg, emg = syn.make_epochs(30, 8192, 64, 64)
eeg_d, emg_d = torch.from_numpy(eeg).to(dev), torch.from_numpy(emg).to(dev)
starts = torch.from_numpy(syn.epoch_segment_starts(30, 8192, 2048, 1024)).to(dev)
tapers = torch.from_numpy(_dpss(2048, 3, 0.9).astype(np.float32)).to(dev)
def timeit(fn, n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): r = fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
print("K1 5 tapers EEG", timeit(lambda: K.fft_segments(eeg_d, starts, tapers, 0, 1, 100)))
print("K1 1 taper EEG", timeit(lambda: K.fft_segments(eeg_d, starts, tapers[:1], 0, 1, 100)))
X = K.fft_segments(eeg_d, starts, tapers, 0, 1, 100); Y = K.fft_segments(emg_d, starts, tapers, 0, 1, 100)
print("K2w JK", timeit(lambda: K.msc_windows(X, Y, None, True, 2.776, 0.81)))
print("K2w noJK", timeit(lambda: K.msc_windows(X, Y, None, False, 0.0, None)))
print("K2w JK maxemg", timeit(lambda: K.msc_windows_maxemg(X, Y, None, True, 2.776, None)))
# mimic bench order: big phase-null workspace first, then the multitaper step
Xp = K.fft_segments(eeg_d, starts, tapers[:1], 1, 1, 100)[:, 0]; Yp = K.fft_segments(emg_d, starts, tapers[:1], 1, 1, 100)[:, 0]
res = K.csd_msc(Xp, Yp)
K.surrogate_null(res, K.SURR_PHASE, 0, 10000, seed=7); torch.cuda.synchronize()
def mt_step():
    Xw = K.fft_segments(eeg_d, starts, tapers, 0, 1, 100)
    Yw = K.fft_segments(emg_d, starts, tapers, 0, 1, 100)
    return K.msc_windows(Xw, Yw, None, True, 2.776, 0.81)
import time
for i in range(5):
    torch.cuda.synchronize(); t0 = time.perf_counter(); o = mt_step(); torch.cuda.synchronize(); print("mt_step wall ms", (time.perf_counter() - t0) * 1e3)
print(torch.cuda.memory_allocated() / 1e9, torch.cuda.memory_reserved() / 1e9)
