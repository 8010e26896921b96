#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the UNMODIFIED reference functions.

Run in the build container only (needs /root/reference):
    python scripts/make_golden.py
The fixtures are committed; the GPU box and the CPU test-suite only read them.
Inputs are stored next to the outputs so no RNG stream has to be reproduced.
"""
from __future__ import annotations

import os
import sys

import numpy as np
from scipy import signal

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.ref_loader import load_reference  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def _signals(seed, n, ne, nm, fs):
    rng = np.random.default_rng(seed)
    t = np.arange(n) / fs
    src = np.sin(2 * np.pi * 20.0 * t + 0.3) + 0.5 * rng.standard_normal(n)
    eeg = rng.standard_normal((n, ne)) + np.outer(src, np.linspace(0.0, 0.9, ne))
    emg = rng.standard_normal((n, nm)) + np.outer(np.roll(src, 3), np.linspace(0.8, 0.0, nm))
    # float32-representable values: the CUDA path ingests float32
    return eeg.astype(np.float32).astype(np.float64), emg.astype(np.float32).astype(np.float64)


def main():
    sf, ds = load_reference()
    os.makedirs(OUT, exist_ok=True)

    # ---- multitaper MSC, no jackknife, independence threshold on ----
    fs = 256.0
    eeg, emg = _signals(11, 1500, 3, 4, fs)
    r = sf.multitaper_magnitude_squared_coherence(
        eeg, emg, fs, nw=3, window_length_sec=1.0, overlap_frac=0.5, use_jackknife=False,
        apply_independence_threshold=True, significance_level=0.05)
    np.savez_compressed(
        os.path.join(OUT, "msc_nojk.npz"), eeg=eeg, emg=emg, fs=fs,
        coherence_raw=r["coherence_raw"], coherence_significant=r["coherence_significant"],
        time_centers=r["time_centers"], freqs=r["freqs"],
        K=r["metadata"]["K_tapers"], IT=r["metadata"]["IT_unadjusted"],
        n_significant=r["metadata"]["n_significant"])

    # ---- multitaper MSC with jackknife CI, window mask, Bonferroni ----
    eeg, emg = _signals(12, 1400, 4, 3, fs)
    W = (1400 - 256) // 128 + 1
    mask = np.ones(W, dtype=bool)
    mask[[1, 4]] = False
    r = sf.multitaper_magnitude_squared_coherence(
        eeg, emg, fs, nw=3, window_length_sec=1.0, overlap_frac=0.5, use_jackknife=True,
        jackknife_alpha=0.05, apply_independence_threshold=True,
        apply_bonferroni_correction=True, significance_level=0.2, window_mask=mask)
    np.savez_compressed(
        os.path.join(OUT, "msc_jk.npz"), eeg=eeg, emg=emg, fs=fs, window_mask=mask,
        coherence_raw=r["coherence_raw"], ci_lower=r["coherence_ci_lower"],
        ci_upper=r["coherence_ci_upper"], coherence_significant=r["coherence_significant"],
        time_centers=r["time_centers"], freqs=r["freqs"],
        IT_unadjusted=r["metadata"]["IT_unadjusted"], IT_bonferroni=r["metadata"]["IT_bonferroni"])

    # ---- jackknife helper called directly on one window (nw=4 -> 7 tapers) ----
    eeg, emg = _signals(13, 128, 2, 3, fs)
    tapers, eigs = signal.windows.dpss(M=128, NW=4, Kmax=7, return_ratios=True)
    tn = [t / np.sqrt(np.sum(t ** 2)) for t in tapers[eigs > 0.9]]
    m, lo, hi = sf.jackknife_coherence_and_ci(tn, eeg, emg, fs, 128, jackknife_alpha=0.1)
    np.savez_compressed(os.path.join(OUT, "jackknife_window.npz"), eeg=eeg, emg=emg, fs=fs,
                        tapers=np.stack(tn), mean=m, lower=lo, upper=hi)

    # ---- different overlap / eeg_axis=1 layout ----
    eeg, emg = _signals(14, 1100, 2, 2, fs)
    r = sf.multitaper_magnitude_squared_coherence(
        eeg.T.copy(), emg, fs, nw=3, window_length_sec=0.5, overlap_frac=0.75, eeg_axis=1,
        use_jackknife=False, apply_independence_threshold=False)
    np.savez_compressed(os.path.join(OUT, "msc_axis_overlap.npz"), eeg=eeg, emg=emg, fs=fs,
                        coherence_raw=r["coherence_raw"], time_centers=r["time_centers"],
                        freqs=r["freqs"])

    # ---- EMG-argmax reduction ----
    rng = np.random.default_rng(15)
    c = rng.random((5, 9, 3, 6)).astype(np.float32)
    lo = (c * 0.5).astype(np.float32)
    hi = np.minimum(c * 1.5, 1).astype(np.float32)
    a, b, d = sf.max_cmc_spectrograms_over_channels(c, lo, hi, channel_ax=3, verbose=False)
    np.savez_compressed(os.path.join(OUT, "max_over_emg.npz"), c=c, lo=lo, hi=hi, a=a, b=b, d=d)

    # ---- multitaper PSD (log and linear) ----
    rng = np.random.default_rng(16)
    x = rng.standard_normal((1024 + 3 * 64, 3)).astype(np.float32).astype(np.float64) + 0.25
    s_log, tc, fr = sf.multitaper_psd(x, fs, nw=3, window_length_sec=0.5, overlap_frac=0.5, axis=0,
                                      apply_log_scale=True)
    s_lin, _, _ = sf.multitaper_psd(x, fs, nw=3, window_length_sec=0.5, overlap_frac=0.5, axis=0,
                                    apply_log_scale=False)
    np.savez_compressed(os.path.join(OUT, "psd.npz"), x=x, fs=fs, s_log=s_log, s_lin=s_lin,
                        time_centers=tc, freqs=fr)

    # ---- Welch MSC = scipy.signal.coherence (preprocessing.py:1228-1230 usage) ----
    eeg, emg = _signals(17, 6000, 2, 3, fs)
    coh = np.zeros((129, 2, 3))
    for i in range(2):
        for j in range(3):
            f, coh[:, i, j] = signal.coherence(eeg[:, i], emg[:, j], fs=fs, nperseg=256)
    np.savez_compressed(os.path.join(OUT, "welch.npz"), eeg=eeg, emg=emg, fs=fs, freqs=f, coh=coh)

    # ---- band / channel aggregation of stored spectrograms (host glue, N1) ----
    rng = np.random.default_rng(18)
    spec = rng.random((6, 129, 5))
    fr = np.fft.rfftfreq(256, 1 / 256.0)
    lo_a, hi_a = spec * 0.8, spec * 1.2
    bands = {'alpha': (8, 12), 'beta': (13, 30)}
    out = {}
    for beh in ('mean', 'max'):
        d = sf.aggregate_spectrogram_over_frequency_band(spec, fr, behaviour=beh, frequency_bands=bands,
                                                         lower_array=lo_a, upper_array=hi_a)
        for b in bands:
            for k, nm in enumerate(('v', 'lo', 'hi')):
                out[f"band_{beh}_{b}_{nm}"] = d[b][k]
    d = sf.aggregate_spectrogram_over_frequency_band(spec, fr, behaviour='mean', frequency_bands=bands,
                                                     log_transform=True, pre_aggregate_axis=(2, 'max'))
    out["band_pre_alpha"] = d['alpha']
    out["psd_agg_emg"] = sf.aggregate_psd_spectrogram(spec, fr, normalize_mvc=True, freq_slice='slow',
                                                      aggregation_ops=[('mean', 1), ('max', 1)])
    out["psd_agg_eeg"] = sf.aggregate_psd_spectrogram(spec, fr, channel_indices=[0, 1, 4], freq_slice=(8, 12),
                                                      aggregation_ops=[('mean', 2), ('mean', 1)])
    np.savez_compressed(os.path.join(OUT, "aggregation.npz"), spec=spec, freqs=fr, **out)

    # ---- force-cycle phase normalisation (CBPA front-end, N2) ----
    import src.pipeline.data_analysis as da
    rng = np.random.default_rng(21)
    t = np.sort(rng.uniform(0, 30, 150))
    t[10] = t[11]                                            # duplicate time stamp
    sig1 = np.sin(2 * np.pi * 0.1 * t) + 0.1 * rng.standard_normal(len(t))
    sig2 = np.stack([sig1, np.cos(2 * np.pi * 0.1 * t), rng.standard_normal(len(t))], axis=1)
    grid36 = np.linspace(0, 360, 36, endpoint=False)
    gridc = np.linspace(0, 360, 13)
    specs = [
        ("a", sig1, t, 0.1, 30.0, grid36, 2, dict(start_offset_sec=10.0, verbose=False)),
        ("b", sig2, t, 0.1, 30.0, grid36, 2, dict(start_offset_sec=0.0, min_cycle_coverage_ratio=0.5, verbose=False)),
        ("c", sig2, t, 0.2, 29.0, gridc, 3, dict(interpolation_kind='nearest', verbose=False)),
        ("d", sig2, t, 0.1, 30.0, grid36, 2, dict(use_interpolation=False, verbose=False)),
        ("e", sig1[:40], t[:40], 0.15, 9.0, gridc, 2,
         dict(min_cycle_coverage_ratio=0.0, phase_wraparound_coverage_threshold=0.95, verbose=False)),
    ]
    pn = {"t": t, "sig1": sig1, "sig2": sig2, "grid36": grid36, "gridc": gridc}
    for name, sgl, tt, f0, dur, g, m, kw in specs:
        cyc = da.phase_normalize_cycles(sgl, tt, f0, dur, g, m, **kw)
        pn[f"n_{name}"] = len(cyc)
        pn[f"cyc_{name}"] = np.stack(cyc) if len(cyc) else np.zeros(0)
    np.savez_compressed(os.path.join(OUT, "phase_norm.npz"), **pn)

    # ---- scalars ----
    np.savez_compressed(
        os.path.join(OUT, "scalars.npz"),
        IT_K5_a05=sf.compute_cmc_independence_threshold(5, 0.05),
        IT_K7_a01=sf.compute_cmc_independence_threshold(7, 0.01),
        fisher_in=np.linspace(0, 1, 11), fisher_out=sf.fisher_atanh_transform(np.linspace(0, 1, 11)),
        inv_in=np.linspace(-3, 3, 13), inv_out=sf.inverse_fisher_atanh(np.linspace(-3, 3, 13)))
    contrast_golden()
    print("golden vectors written to", OUT)
    for fn in sorted(os.listdir(OUT)):
        print(f"  {fn:28s} {os.path.getsize(os.path.join(OUT, fn)) / 1024:8.1f} KiB")


def contrast_golden():
    """X / channel names / grid of the reference's ``cbpa.build_contrast_array`` (cbpa.py:733-942) on the synthetic
    study of tests/cbpa_fixture.py.  The reference module imports mne and matplotlib at the top; neither is touched
    by the contrast front-end, so they are replaced by inert stubs for the import."""
    import tempfile
    import warnings
    from unittest.mock import MagicMock
    for name in ("mne", "mne.stats", "mne.channels", "mne.io", "matplotlib", "matplotlib.pyplot",
                 "src.pipeline.visualizations"):
        sys.modules.setdefault(name, MagicMock())
    import src.pipeline.cbpa as ref_cbpa          # noqa: E402  (load_reference() put /root/reference on sys.path)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import cbpa_fixture as fx                      # noqa: E402
    out = {}
    with tempfile.TemporaryDirectory() as tmp:
        root = fx.build_study(os.path.join(tmp, "study"))
        for name, cfg in fx.configs(ref_cbpa, root, os.path.join(tmp, "out")).items():
            with warnings.catch_warnings(record=True) as w:
                warnings.simplefilter("always")
                X, ch, grid = ref_cbpa.build_contrast_array(cfg)
            out[f"X_{name}"] = X
            out[f"ch_{name}"] = np.array(list(ch))
            out[f"grid_{name}"] = np.asarray(grid)
            out[f"n_warnings_{name}"] = len([x for x in w if "Skipping" in str(x.message)])
            print(f"  contrast {name}: X {X.shape}, {out[f'n_warnings_{name}']} skip warnings")
    np.savez_compressed(os.path.join(OUT, "contrast.npz"), **out)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "contrast":
        load_reference()
        contrast_golden()
    else:
        main()
