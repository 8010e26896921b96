"""Per-role timeline of CTA 0 of the tensor-core Welch kernel (CMC_DT_DBG bit 32; globaltimer stamps)."""
import ctypes as C
import os
import sys

os.environ["CMC_DT_DBG"] = str(int(os.environ.get("CMC_DT_DBG", "0")) | 32)
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from multimodal_biosignal_analysis_b200 import _lib, kernels as K, synthetic as syn

dev = torch.device("cuda:0")
NE = int(os.environ.get("NE", "30"))
eeg, emg = syn.make_epochs(NE, 8192, 64, 64, seed=20260102)
st = syn.epoch_segment_starts(NE, 8192, 2048, 1024)
e_d, m_d = torch.from_numpy(eeg).to(dev), torch.from_numpy(emg).to(dev)
plan = K.WelchHannPlan(st, 2048, 1, 100)
spec = torch.empty((len(st), 1, 100, 128), dtype=torch.complex64, device=dev)
for _ in range(3):
    plan.spectra(e_d, spec[..., :64], m_d, spec[..., 64:])
torch.cuda.synchronize()
buf = (C.c_ulonglong * 64)()
lib = _lib.load()
lib.cmc_dbg_dt_trace.argtypes = [C.c_void_p]
lib.cmc_dbg_dt_trace(buf)
t = [int(v) for v in buf]
t0 = t[0]
rel = lambda i: (t[i] - t0) / 1e3
print(f"epochs {NE}: prologue done {rel(1):.2f} us, epilogue start {rel(50):.2f}, epilogue end {rel(51):.2f}, kernel end {rel(52):.2f}")
print("kb   tma-issue  landed   mma-issue")
for kb in range(16):
    print(f"{kb:2d}   {rel(2 + kb):8.2f} {rel(34 + kb):8.2f} {rel(18 + kb):8.2f}")
print("converter (kb 4-6): landed -> loads done -> planes free -> planes written")
for i in range(3):
    print(f"{4 + i:2d}   {rel(34 + 4 + i):8.2f} {rel(53 + i):8.2f} {rel(56 + i):8.2f} {rel(59 + i):8.2f}")
