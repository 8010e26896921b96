"""One small call of every kernel family, meant to run under compute-sanitizer (scripts/run_sanitizers.sh):
K1 (pipelined TMA kernel, plain TMA kernel, LDG kernel, direct DFT), K2 (direct MN-major kernel and packed path), K2w,
K3 (shift4, phase GEMM resident / streamed, phase histogram), K4 (observed + permutations) and the PSD epilogue."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from scipy import signal
from scipy.stats import t as t_dist

from multimodal_biosignal_analysis_b200 import cbpa as cb, kernels as K, synthetic as syn

only = set(sys.argv[1:])


def want(name):
    return not only or name in only


dev = torch.device("cuda:0")
rng = np.random.default_rng(0)


def d(a, dtype=None):
    return torch.as_tensor(np.ascontiguousarray(a), dtype=dtype).to(dev)


if want("k1"):
    for N, n_ch, n_seg in ((512, 32, 16), (512, 8, 3), (4096, 8, 2), (256, 5, 3), (100, 3, 2)):
        x = d(rng.standard_normal((N * (n_seg + 1) // 2 + N, n_ch)).astype(np.float32))
        starts = d((np.arange(n_seg) * (N // 2)).astype(np.int64))
        win = d(signal.get_window("hann", N).astype(np.float32)[None])
        for lo, hi in ((0, N // 2), (1, min(40, N // 2 - 1))):
            s = K.fft_segments(x, starts, win, 1, lo, hi)
            assert torch.isfinite(torch.view_as_real(s)).all()
    print("k1 ok")

if want("k1t"):
    # tensor-core Welch kernel: 12 + 60 channels (two partly empty 64-channel groups), mixed chains, both emission phases
    e = d(rng.standard_normal((6000, 12)).astype(np.float32))
    m = d(rng.standard_normal((6000, 60)).astype(np.float32))
    st = np.array([0, 256, 512, 3000, 5400], dtype=np.int64)
    plan = K.WelchHannPlan(st, 512, 1, 40)
    out = torch.empty((len(st), 1, 40, 72), dtype=torch.complex64, device=dev)
    plan.spectra(e, out[..., :12], m, out[..., 12:])
    assert torch.isfinite(torch.view_as_real(out)).all()
    print("k1t ok")

if want("k2") or want("k3") or want("k2w"):
    L, F, ne, nm = 24, 6, 12, 70
    X = d((rng.standard_normal((L, F, ne)) + 1j * rng.standard_normal((L, F, ne))).astype(np.complex64))
    Y = d((rng.standard_normal((L, F, nm)) + 1j * rng.standard_normal((L, F, nm))).astype(np.complex64))
if want("k2"):
    a = K.csd_msc(X, Y, want_sxy=True)
    b = K.csd_msc(X, Y, want_sxy=True, keep_operands=True)
    assert torch.allclose(a.coh, b.coh, atol=1e-5)
    print("k2 ok")
if want("k3"):
    res = K.csd_msc(X, Y)
    sh = d(rng.integers(1, L, 20).astype(np.int32))
    e, m = K.surrogate_null(res, K.SURR_SHIFT, 0, 20, shifts=sh)
    e, m = K.surrogate_null(res, K.SURR_PHASE, 0, 140, seed=3)                   # split operands, resident panel
    h, bl = K.surrogate_null_hist(res, 0, 140, seed=3, n_bins=64)
    assert int(h.sum()) + int(bl.sum()) <= 140 * F * ne * nm
    L2 = 300
    X2 = d((rng.standard_normal((L2, 3, 4)) + 1j * rng.standard_normal((L2, 3, 4))).astype(np.complex64))
    Y2 = d((rng.standard_normal((L2, 3, 6)) + 1j * rng.standard_normal((L2, 3, 6))).astype(np.complex64))
    r2 = K.csd_msc(X2, Y2)
    K.surrogate_null(r2, K.SURR_PHASE, 0, 130, seed=4)                          # streamed panel
    torch.cuda.synchronize()
    print("k3 ok")
if want("k2w"):
    W, Kt = 4, 5
    Xw = d((rng.standard_normal((W, Kt, F, ne)) + 1j * rng.standard_normal((W, Kt, F, ne))).astype(np.complex64))
    Yw = d((rng.standard_normal((W, Kt, F, nm)) + 1j * rng.standard_normal((W, Kt, F, nm))).astype(np.complex64))
    K.msc_windows(Xw, Yw, None, True, 2.776, 0.81)
    K.msc_windows_maxemg(Xw, Yw, None, True, 2.776, 0.81, True, True)
    K.psd_from_spectra(Xw, 1.0, True, 0, 64, True)
    torch.cuda.synchronize()
    print("k2w ok")
if want("k4"):
    Xc = syn.make_cbpa_contrast(12, 20, 16, seed=2)
    adj = cb.combine_adjacency(20, cb.find_ch_adjacency_from_positions(syn.sensor_positions(64)[:16])).tocsr()
    adj.sort_indices()
    signs = syn.make_sign_table(300, 12, seed=3)
    thr = float(t_dist.ppf(0.975, 11))
    Xd = d(Xc.reshape(12, -1))
    ip, ix = d(adj.indptr.astype(np.int32)), d(adj.indices.astype(np.int32))
    ws = K.cbpa_workspace(Xd)
    K.cbpa_observed(Xd, thr, 0, ip, ix, ws=ws)
    h0 = K.cbpa_permute(Xd, d(signs), 0, 300, thr, 0, ip, ix, ws=ws, tiled=True)
    h1 = K.cbpa_permute(Xd, d(signs), 0, 300, thr, 0, ip, ix)
    assert torch.equal(h0, h1)
    print("k4 ok")
torch.cuda.synchronize()
print("sanitize driver done")
