"""Where the per-pair threshold stage of config 3 spends its time (CUDA events around every piece)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from scipy import signal
from multimodal_biosignal_analysis_b200 import kernels as K, synthetic as syn, data_surrogation as ds

dev = torch.device("cuda:0")
eeg, emg = syn.make_epochs(30, 8192, 64, 64, seed=20260102)
starts_h = syn.epoch_segment_starts(30, 8192, 2048, 1024)
starts = torch.from_numpy(starts_h).to(dev)
win = torch.from_numpy(signal.get_window("hann", 2048).astype(np.float32)[None]).to(dev)
L = len(starts_h)
spec = torch.empty((L, 1, 100, 128), dtype=torch.complex64, device=dev)
K.fft_segments_pair(torch.from_numpy(eeg).to(dev), torch.from_numpy(emg).to(dev), starts, win, 1, 1, 100, spec[..., :64], spec[..., 64:])
res = K.csd_msc(spec[:, 0, :, :64], spec[:, 0, :, 64:])
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000


def timed(label, fn, reps=3):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(reps):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    print(f"{label:42s} device {e0.elapsed_time(e1) / reps:8.3f} ms   wall {(time.perf_counter() - t0) * 1e3 / reps:8.3f} ms")
    return out


timed("exceedance null", lambda: K.surrogate_null(res, K.SURR_PHASE, 0, n, seed=3))
timed("exceedance null (operands kept)", lambda: K.surrogate_null(res, K.SURR_PHASE, 0, n, seed=3, keep_phase_operands=True))
m_ = res.null_mean()
lo_ = (1.5 * m_).contiguous(); sc_ = (128.0 / (16.0 * m_).clamp(min=1e-12)).contiguous()
h, _b = timed("hist pass, windowed, operands reused", lambda: K.surrogate_null_hist(res, 0, n, seed=3, bin_lo=lo_, bin_scale=sc_, keep_operands=True))
timed("hist pass, full range [0,1], reused", lambda: K.surrogate_null_hist(res, 0, n, seed=3, keep_operands=True))
res.phase_ops = None
timed("hist pass, operands regenerated", lambda: K.surrogate_null_hist(res, 0, n, seed=3))
below = torch.zeros((100, 64, 64), dtype=torch.int32, device=dev)
timed("hist_select", lambda: K.hist_select(h, int(0.95 * (n - 1)), below.clone()))
timed("torch.zeros hist", lambda: torch.zeros((100, 64, 64, 128), dtype=torch.int32, device=dev))
timed("null_quantile_thresholds (3 passes)", lambda: ds.null_quantile_thresholds(res, n, 3, 0.95, passes=3))
import types
P = types.SimpleNamespace(device_result=res, coherence=res.coh, freqs=np.arange(100), group=1)
timed("phase_randomised_surrogate_null(thresholds)", lambda: ds.phase_randomised_surrogate_null(P, n, seed=3, thresholds=True))
timed("phase_randomised_surrogate_null(plain)", lambda: ds.phase_randomised_surrogate_null(P, n, seed=3))
