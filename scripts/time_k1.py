import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from scipy import signal
from multimodal_biosignal_analysis_b200 import kernels as K, synthetic as syn, _lib
dev = torch.device("cuda")
sets = [torch.from_numpy(syn.make_epochs(30, 8192, 64, 64, seed=s)[0]).to(dev) for s in range(4)]
starts = torch.from_numpy(syn.epoch_segment_starts(30, 8192, 2048, 1024)).to(dev)
win = torch.from_numpy(signal.get_window("hann", 2048).astype(np.float32)[None]).to(dev)
out = torch.empty((len(starts), 1, 100, 64), dtype=torch.complex64, device=dev)
for i in range(8):
    K.fft_segments(sets[i % 4], starts, win, 1, 1, 100, out=out)
torch.cuda.synchronize()
lib = _lib.load()
if hasattr(lib, "cmc_dbg_k1_cycles"):
    buf = (ctypes.c_ulonglong * 8)()
    lib.cmc_dbg_k1_cycles(buf, 1)
n = 40
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(n):
    K.fft_segments(sets[i % 4], starts, win, 1, 1, 100, out=out)
e1.record(); torch.cuda.synchronize()
print("K1 us/launch", e0.elapsed_time(e1) / n * 1e3)
if hasattr(lib, "cmc_dbg_k1_cycles"):
    lib.cmc_dbg_k1_cycles(buf, 0)
    tiles = 210 * 8 / 2 * n            # tiles of worker 0 over all CTAs (about half of all tiles)
    names = ["wait tile", "load regs + mean", "window + radix-16 + store", "pass 2", "pass 3", "split + store"]
    tot = sum(buf[:6])
    for nm, b in zip(names, buf[:6]):
        print(f"{nm:28s} {b / tiles:9.0f} cycles/tile  {b / tot:6.3f}")
