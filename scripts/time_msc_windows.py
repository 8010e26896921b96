"""K2w alone (cmc_msc_windows / cmc_msc_windows_maxemg, jackknife CI) on the multitaper variant of config 2:
210 windows x K = 5 tapers x F = 100 bins x 64 x 64 pairs, random spectra, CUDA events over graph-free launches.
Usage (GPU box): python scripts/time_msc_windows.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from multimodal_biosignal_analysis_b200 import kernels as K

dev = torch.device("cuda:0")
torch.manual_seed(0)
for (W, Kt, F, Ne, Nm) in ((210, 5, 100, 64, 64), (599, 5, 200, 11, 64), (210, 7, 100, 64, 64), (210, 5, 100, 63, 61)):
    X = torch.randn(W, Kt, F, Ne, dtype=torch.complex64, device=dev)
    Y = torch.randn(W, Kt, F, Nm, dtype=torch.complex64, device=dev) + 0.5 * X[..., :1]
    for name, fn in (("msc_windows", lambda: K.msc_windows(X, Y, None, True, 2.776, 0.3)),
                     ("msc_windows_maxemg", lambda: K.msc_windows_maxemg(X, Y, None, True, 2.776, 0.3, True, True))):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        print(f"W={W} K={Kt} F={F} {Ne}x{Nm} {name}: {ms:.3f} ms  ({W * F * Ne * Nm / ms / 1e6:.1f} G pair-outputs/s)")
