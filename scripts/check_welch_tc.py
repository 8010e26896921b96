"""Tensor-core Welch spectra (cmc_welch_hann_*) against the FFT path and the fp64 oracle, then timing.
Usage (GPU box): python scripts/check_welch_tc.py [--time]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from scipy import signal

from multimodal_biosignal_analysis_b200 import kernels as K, synthetic as syn
from oracle import coherence as oc

dev = torch.device("cuda:0")
FAIL = []


def oracle_spec(x, starts, N, detrend, lo, hi):
    """fp64 hann-windowed spectra (n_seg, F, C) with the detrend conventions of cmc_fft_segments."""
    win = signal.get_window("hann", N)
    out = np.empty((len(starts), hi - lo + 1, x.shape[1]), dtype=np.complex128)
    for i, s in enumerate(starts):
        seg = x[s:s + N].astype(np.float64)
        if detrend == K.DETREND_CONSTANT:
            seg = seg - seg.mean(axis=0, keepdims=True)
        sp = np.fft.rfft(seg * win[:, None], axis=0)
        if detrend == K.DETREND_POST_TAPER:
            sp[0] = 0.0
        out[i] = sp[lo:hi + 1]
    return out


def case(name, eeg, emg, starts, N, lo, hi, detrend, tol=6e-5):
    starts = np.asarray(starts, dtype=np.int64)
    F = hi - lo + 1
    ne = eeg.shape[1]
    nm = 0 if emg is None else emg.shape[1]
    plan = K.WelchHannPlan(starts, N, lo, hi)
    e_d = torch.from_numpy(eeg).to(dev)
    m_d = None if emg is None else torch.from_numpy(emg).to(dev)
    spec = torch.full((len(starts), 1, F, ne + nm), float("nan"), dtype=torch.complex64, device=dev)
    if emg is None:
        plan.spectra(e_d, spec)
    else:
        plan.spectra(e_d, spec[..., :ne], m_d, spec[..., ne:], detrend=detrend)
    if emg is None and detrend != K.DETREND_CONSTANT:
        raise SystemExit("single-recording case uses constant detrend")
    torch.cuda.synchronize()
    got = spec[:, 0].cpu().numpy()
    ref = oracle_spec(np.concatenate([eeg] + ([emg] if emg is not None else []), axis=1), starts, N, detrend, lo, hi)
    # FFT path for comparison of the error level
    win = torch.from_numpy(signal.get_window("hann", N).astype(np.float32)[None]).to(dev)
    sd = torch.from_numpy(starts).to(dev)
    fft = torch.empty_like(spec)
    K.fft_segments(e_d, sd, win, detrend, lo, hi, out=fft, ch_offset=0)
    if emg is not None:
        K.fft_segments(m_d, sd, win, detrend, lo, hi, out=fft, ch_offset=ne)
    torch.cuda.synchronize()
    fftn = fft[:, 0].cpu().numpy()
    # error relative to the rms of the spectrum of each channel (what a coherence sees)
    rms = np.sqrt(np.mean(np.abs(ref) ** 2, axis=(0, 1), keepdims=True)) + 1e-30
    e_tc = float(np.nanmax(np.abs(got - ref) / rms))
    e_fft = float(np.nanmax(np.abs(fftn - ref) / rms))
    nan = int(np.isnan(got.view(np.float32)).sum())
    ok = nan == 0 and e_tc < tol
    print(f"{'ok  ' if ok else 'FAIL'} {name}: half blocks {plan.n_half_blocks}, max |dX| / rms  tc {e_tc:.2e}  fft {e_fft:.2e}  nan {nan}",
          flush=True)
    if not ok:
        FAIL.append(name)
    # run-to-run determinism
    spec2 = torch.empty_like(spec)
    if emg is None:
        plan.spectra(e_d, spec2)
    else:
        plan.spectra(e_d, spec2[..., :ne], m_d, spec2[..., ne:], detrend=detrend)
    torch.cuda.synchronize()
    if not torch.equal(spec2.view(torch.float32), spec.view(torch.float32)):
        print(f"FAIL {name}: not bit-identical between two runs")
        FAIL.append(name + " determinism")
    return got, ref


def main():
    rng = np.random.default_rng(0)
    # 1. small: 3 epochs of config 2
    eeg, emg = syn.make_epochs(3, 8192, 64, 64, seed=1)
    st = syn.epoch_segment_starts(3, 8192, 2048, 1024)
    case("cfg2 x 3 epochs, constant", eeg, emg, st, 2048, 1, 100, K.DETREND_CONSTANT)
    case("cfg2 x 3 epochs, none", eeg, emg, st, 2048, 1, 100, K.DETREND_NONE)
    case("cfg2 x 3 epochs, post-taper, bins 0-99", eeg, emg, st, 2048, 0, 99, K.DETREND_POST_TAPER)
    case("bins 0-101 constant", eeg, emg, st, 2048, 0, 101, K.DETREND_CONSTANT)
    case("bins 40-141", eeg, emg, st, 2048, 40, 141, K.DETREND_CONSTANT)
    # 2. DC offset 1000 x the signal and a slow drift
    t = np.arange(eeg.shape[0], dtype=np.float32)[:, None]
    off_e = eeg + 1000.0 + 0.01 * t
    off_m = emg - 300.0
    case("dc offset + drift", off_e.astype(np.float32), off_m.astype(np.float32), st, 2048, 1, 100, K.DETREND_CONSTANT, tol=1e-3)  # float32 input quantisation: the FFT kernel sits at 9e-4
    # 3. odd shapes: 12 x 60 channels, shuffled isolated / chained segments, N = 1024
    e2 = rng.standard_normal((20000, 12)).astype(np.float32)
    m2 = rng.standard_normal((20000, 60)).astype(np.float32)
    st2 = np.array([0, 512, 1024, 5000, 5512, 9000, 3000, 3512, 4024, 18976], dtype=np.int64)
    case("12 x 60 ch, N=1024, mixed chains", e2, m2, st2, 1024, 1, 50, K.DETREND_CONSTANT)
    # 4. one recording of 128 channels, and one of 200 (two group pairs)
    x3 = rng.standard_normal((30000, 128)).astype(np.float32)
    case("single 128 ch, N=512", x3, None, np.arange(0, 30000 - 512, 256), 512, 2, 60, K.DETREND_CONSTANT)
    x4 = rng.standard_normal((9000, 200)).astype(np.float32)
    case("single 200 ch, N=4096", x4, None, np.arange(0, 9000 - 4096, 2048), 4096, 1, 100, K.DETREND_CONSTANT)
    # 5. coloured noise (1/f^2 power): weak high bins next to strong low ones
    w = rng.standard_normal((40000, 64)).astype(np.float64)
    col = np.cumsum(w, axis=0)
    col -= col.mean(axis=0)
    case("brownian noise", col.astype(np.float32), (col[:, ::-1] * 0.5 + w).astype(np.float32),
         np.arange(0, 40000 - 2048, 1024), 2048, 1, 100, K.DETREND_CONSTANT, tol=1e-3)
    # 6. full config 2 and the coherence it gives
    eeg, emg = syn.make_epochs(30, 8192, 64, 64, seed=20260102)
    st = syn.epoch_segment_starts(30, 8192, 2048, 1024)
    got, ref = case("config 2 full", eeg, emg, st, 2048, 1, 100, K.DETREND_CONSTANT)
    c_ref = oc.msc_from_spectra(ref[:, :, :64], ref[:, :, 64:])[0]
    c_got = oc.msc_from_spectra(got[:, :, :64].astype(np.complex128), got[:, :, 64:].astype(np.complex128))[0]
    dc = float(np.max(np.abs(c_ref - c_got)))
    print(f"{'ok  ' if dc < 1e-4 else 'FAIL'} coherence from tc spectra vs fp64: max |dC| {dc:.2e}")
    if dc >= 1e-4:
        FAIL.append("coherence")

    if "--time" in sys.argv:
        e_d, m_d = torch.from_numpy(eeg).to(dev), torch.from_numpy(emg).to(dev)
        e_b, m_b = torch.from_numpy(eeg[::-1].copy()).to(dev), torch.from_numpy(emg[::-1].copy()).to(dev)
        plan = K.WelchHannPlan(st, 2048, 1, 100)
        sd = torch.from_numpy(st).to(dev)
        win = torch.from_numpy(signal.get_window("hann", 2048).astype(np.float32)[None]).to(dev)
        spec = torch.empty((len(st), 1, 100, 128), dtype=torch.complex64, device=dev)
        for label, fn in (("tensor-core", lambda a, b: plan.spectra(a, spec[..., :64], b, spec[..., 64:])),
                          ("fft pair", lambda a, b: K.fft_segments_pair(a, b, sd, win, K.DETREND_CONSTANT, 1, 100,
                                                                        spec[..., :64], spec[..., 64:]))):
            for _ in range(5):
                fn(e_d, m_d)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(50):
                fn(*((e_d, m_d) if i & 1 else (e_b, m_b)))
            e1.record()
            torch.cuda.synchronize()
            print(f"{label}: {e0.elapsed_time(e1) / 50 * 1e3:.1f} us per launch (config 2, both modalities)")
    print("FAILED: " + ", ".join(FAIL) if FAIL else "all ok")
    return 1 if FAIL else 0


if __name__ == "__main__":
    sys.exit(main())
