"""Every kernel of the hot path once at its BASELINE size, for `ncu --set full` (scripts/profile_r2.sh):
config 2 (K1t and the FFT K1 for both modalities, K2 direct), config 3 (shift null, phase null of 1,024 surrogates, per-pair histogram pass),
the per-window multitaper estimator (K2w, 210 windows x 5 tapers), config 4 CBPA (1,184 permutations = 4 per CTA)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from scipy import signal
from scipy.stats import t as t_dist

from multimodal_biosignal_analysis_b200 import cbpa as cb, kernels as K, synthetic as syn
from multimodal_biosignal_analysis_b200.signal_features import _dpss

dev = torch.device("cuda:0")
eeg, emg = syn.make_epochs(30, 8192, 64, 64, seed=20260102)
starts_h = syn.epoch_segment_starts(30, 8192, 2048, 1024)
eeg_d, emg_d = torch.from_numpy(eeg).to(dev), torch.from_numpy(emg).to(dev)
starts = torch.from_numpy(starts_h).to(dev)
win = torch.from_numpy(signal.get_window("hann", 2048).astype(np.float32)[None]).to(dev)
L = len(starts_h)
spec = torch.empty((L, 1, 100, 128), dtype=torch.complex64, device=dev)
plan = K.WelchHannPlan(starts_h, 2048, 1, 100)     # K1t: the tensor-core Welch kernel (default K1 of the Welch path)
for _ in range(2):                                  # first pass warms tables / attributes, second one is profiled
    plan.spectra(eeg_d, spec[..., :64], emg_d, spec[..., 64:])
    K.fft_segments_pair(eeg_d, emg_d, starts, win, K.DETREND_CONSTANT, 1, 100, spec[..., :64], spec[..., 64:])
    res = K.csd_msc(spec[:, 0, :, :64], spec[:, 0, :, 64:])
    shifts = torch.from_numpy(np.random.default_rng(3).integers(1, L, 1000).astype(np.int32)).to(dev)
    K.surrogate_null(res, K.SURR_SHIFT, 0, 1000, shifts=shifts)
    K.surrogate_null(res, K.SURR_PHASE, 0, 1024, seed=7)
    m_ = res.null_mean()                            # the first window of the per-pair threshold search
    K.surrogate_null_hist(res, 0, 1024, seed=7, n_bins=128, bin_lo=(1.5 * m_).contiguous(),
                          bin_scale=(128.0 / (16.0 * m_).clamp(min=1e-12)).contiguous())
    tapers = torch.from_numpy(_dpss(2048, 3, 0.9).astype(np.float32)).to(dev)
    S = torch.empty((L, tapers.shape[0], 100, 128), dtype=torch.complex64, device=dev)
    K.fft_segments_pair(eeg_d, emg_d, starts, tapers, K.DETREND_NONE, 1, 100, S[..., :64], S[..., 64:])
    K.msc_windows(S[..., :64], S[..., 64:], None, True, float(t_dist.ppf(0.975, 4)), 0.81)
    K.msc_windows_maxemg(S[..., :64], S[..., 64:], None, True, float(t_dist.ppf(0.975, 4)), 0.81, True, False)
    del S
    adj = cb.combine_adjacency(100, cb.find_ch_adjacency_from_positions(syn.sensor_positions(64))).tocsr()
    adj.sort_indices()
    Xc = torch.from_numpy(np.ascontiguousarray(syn.make_cbpa_contrast(20, 100, 64).reshape(20, -1))).to(dev)
    sd = torch.from_numpy(syn.make_sign_table(1184, 20, seed=42)).to(dev)
    ip = torch.from_numpy(adj.indptr.astype(np.int32)).to(dev)
    ix = torch.from_numpy(adj.indices.astype(np.int32)).to(dev)
    ws = K.cbpa_workspace(Xc)
    K.cbpa_observed(Xc, float(t_dist.ppf(0.975, 19)), 0, ip, ix, ws=ws)
    K.cbpa_permute(Xc, sd, 0, 1184, float(t_dist.ppf(0.975, 19)), 0, ip, ix, ws=ws, tiled=True)
    torch.cuda.synchronize()
print("profile driver done")
