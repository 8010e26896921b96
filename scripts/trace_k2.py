"""Timeline of one csd_mn_kernel tile (CTA 0), instrumented build: CMC_NVCC_EXTRA=-DCMC_K2_TRACE."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from scipy import signal
from multimodal_biosignal_analysis_b200 import kernels as K, synthetic as syn, _lib
dev = torch.device("cuda")
eeg, emg = syn.make_epochs(30, 8192, 64, 64)
starts = torch.from_numpy(syn.epoch_segment_starts(30, 8192, 2048, 1024)).to(dev)
win = torch.from_numpy(signal.get_window("hann", 2048).astype(np.float32)[None]).to(dev)
spec = torch.empty((len(starts), 1, 100, 128), dtype=torch.complex64, device=dev)
K.fft_segments(torch.from_numpy(eeg).to(dev), starts, win, 1, 1, 100, out=spec, ch_offset=0)
K.fft_segments(torch.from_numpy(emg).to(dev), starts, win, 1, 1, 100, out=spec, ch_offset=64)
for _ in range(5):
    res = K.csd_msc(spec[:, 0, :, :64], spec[:, 0, :, 64:])
torch.cuda.synchronize()
lib = _lib.load()
buf = (ctypes.c_ulonglong * 64)()
lib.cmc_dbg_k2_trace(buf)
t = np.array(buf[:], dtype=np.int64)
t0 = t[0]
names = {0: "start", 1: "setup done", 40: "epi: tmem_full", 41: "epi: tmem read+fold done", 42: "epi: stores done", 43: "exit"}
for k in range(7):
    names[2 + k] = f"tma issue kb{k}"; names[10 + k] = f"conv sees full kb{k}"; names[20 + k] = f"conv done kb{k}"; names[30 + k] = f"mma sees conv kb{k}"
for slot in sorted(names, key=lambda s: t[s]):
    if t[slot]:
        print(f"{(t[slot] - t0) / 1e3:8.2f} us  {names[slot]}")
