"""numpy (fp64) restatement of the reference's coherence / PSD arithmetic.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  All line citations are into
``src/pipeline/signal_features.py`` of the reference unless stated otherwise.

The restatement is vectorised over windows (the reference loops) but performs
the same arithmetic: taper -> rFFT -> sum over tapers of conj(X) Y and |X|^2 ->
|S_xy|^2 / max(S_xx S_yy, tiny) clipped to [0, 1].  Pinned against the
reference's own output in ``tests/test_oracle_golden.py``.
"""
from __future__ import annotations

import numpy as np
from scipy import signal
from scipy.stats import beta as _beta, t as _t_dist

TINY = np.finfo(np.float64).tiny


# --------------------------------------------------------------------------
# window grid and tapers
# --------------------------------------------------------------------------
def window_params(sampling_freq: float, window_length_sec: float, overlap_frac: float):
    """(window_samples, hop_samples) exactly as :667-668."""
    window_samples = int(window_length_sec * sampling_freq)
    hop_samples = int(window_samples * (1 - overlap_frac))
    return window_samples, hop_samples


def n_windows_msc(n_samples: int, window_samples: int, hop_samples: int) -> int:
    """Window count of the MSC grid, :682."""
    return (n_samples - window_samples) // hop_samples + 1


def window_starts_psd(n_samples: int, window_samples: int, hop_samples: int) -> np.ndarray:
    """Window starts of ``multitaper_psd`` (:398) - one window fewer than the
    MSC grid when (n - N) % hop == 0."""
    return np.arange(0, n_samples - window_samples, hop_samples)


def dpss_tapers(window_samples: int, nw: float = 3, eig_threshold: float = 0.90,
                normalise: bool = True):
    """DPSS tapers kept by eigenvalue and L2-normalised, :669-678."""
    k = int(2 * nw - 1)
    tapers, eigs = signal.windows.dpss(M=window_samples, NW=nw, Kmax=k, return_ratios=True)
    keep = eigs > eig_threshold
    tapers = tapers[keep]
    if normalise:
        tapers = np.stack([t / np.sqrt(np.sum(t ** 2)) for t in tapers])
    return tapers, eigs


# --------------------------------------------------------------------------
# spectra of segments
# --------------------------------------------------------------------------
def segment_spectra(x: np.ndarray, starts: np.ndarray, windows: np.ndarray,
                    detrend: int = 0, bin_lo: int = 0, bin_hi: int | None = None) -> np.ndarray:
    """rFFT of every (segment, window-row, channel).

    x        (n_samples, n_ch) real
    starts   (n_seg,) first sample of each segment
    windows  (K, N) taper / window rows
    detrend  0 none (multitaper MSC, :743-748)
             1 remove segment mean BEFORE windowing (scipy Welch detrend='constant')
             2 remove mean of the tapered segment AFTER windowing
               (scipy.signal.periodogram(window=None) default detrend, :419)
    returns  complex128 (n_seg, K, F, n_ch), F = bin_hi - bin_lo + 1
    """
    x = np.asarray(x, dtype=np.float64)
    windows = np.asarray(windows, dtype=np.float64)
    K, N = windows.shape
    idx = np.asarray(starts)[:, None] + np.arange(N)[None, :]
    seg = x[idx]                                   # (n_seg, N, n_ch)
    if detrend == 1:
        seg = seg - seg.mean(axis=1, keepdims=True)
    tap = seg[:, None, :, :] * windows[None, :, :, None]      # (n_seg, K, N, n_ch)
    if detrend == 2:
        tap = tap - tap.mean(axis=2, keepdims=True)
    spec = np.fft.rfft(tap, axis=2)
    if bin_hi is None:
        bin_hi = N // 2
    return spec[:, :, bin_lo:bin_hi + 1, :]


def msc_from_spectra(X: np.ndarray, Y: np.ndarray):
    """Pooled magnitude-squared coherence over the leading axis.

    X (L, F, Ne), Y (L, F, Nm) complex.  Restates :750-770 (the 1/(fs N) and
    1/K factors cancel in the ratio and are dropped):
        C = clip(|sum conj(X) Y|^2 / max(sum|X|^2 * sum|Y|^2, tiny), 0, 1)
    returns (coh (F,Ne,Nm), sxx (F,Ne), syy (F,Nm), sxy (F,Ne,Nm) complex)
    """
    sxx = np.sum(np.abs(X) ** 2, axis=0)
    syy = np.sum(np.abs(Y) ** 2, axis=0)
    sxy = np.einsum("lfi,lfj->fij", np.conj(X), Y)
    den = np.maximum(sxx[:, :, None] * syy[:, None, :], TINY)
    coh = np.clip(np.abs(sxy) ** 2 / den, 0.0, 1.0)
    return coh, sxx, syy, sxy


# --------------------------------------------------------------------------
# Fisher pair, independence threshold  (:459-481, :581-604)
# --------------------------------------------------------------------------
def fisher_atanh_transform(c, eps: float = 1e-10):
    cs = np.clip(c, eps, 1 - eps)
    return 0.5 * np.log((1 + cs) / (1 - cs))


def inverse_fisher_atanh(z):
    return np.tanh(z) ** 2


def independence_threshold(K: int, alpha: float = 0.05) -> float:
    return float(_beta.ppf(1 - alpha, K - 2, K - 2))


def threshold_filtering(coh, K, alpha=0.05, n_comparisons=None, apply_bonferroni=False):
    if apply_bonferroni and n_comparisons is not None:
        a = alpha / n_comparisons
        if a < 1e-10:
            a = 1e-10
    else:
        a = alpha
    it = independence_threshold(K, a)
    return coh > it, it


# --------------------------------------------------------------------------
# jackknife  (:484-578)
# --------------------------------------------------------------------------
def jackknife_from_spectra(X: np.ndarray, Y: np.ndarray, alpha: float = 0.05):
    """Leave-one-taper-out coherence mean and Student-t CI for ONE window.

    X (K, F, Ne), Y (K, F, Nm).  The reference recomputes the K-1 FFTs per
    replicate (:507-531); sums over the other tapers are the same numbers, so
    the restatement subtracts taper k from the total.  Mean in coherence space
    (:555-556), variance (K-1)/K sum (z_k - zbar)^2 in Fisher-z space
    (:559-562), CI = tanh(z(mean) -+ t_crit se)^2 forced to bracket the mean
    (:565-576).
    """
    K = X.shape[0]
    pxx = np.abs(X) ** 2                       # (K,F,Ne)
    pyy = np.abs(Y) ** 2
    cxy = np.conj(X)[:, :, :, None] * Y[:, :, None, :]      # (K,F,Ne,Nm)
    sxx = pxx.sum(0)[None] - pxx
    syy = pyy.sum(0)[None] - pyy
    sxy = cxy.sum(0)[None] - cxy
    # the common 1/((K-1) fs N) factors cancel between numerator and denominator
    num = np.abs(sxy) ** 2
    den = sxx[:, :, :, None] * syy[:, :, None, :]
    with np.errstate(divide="ignore", invalid="ignore"):
        coh_k = np.clip(num / np.maximum(den, TINY), 0.0, 1.0)
    z_k = fisher_atanh_transform(coh_k)
    mean = np.clip(coh_k.mean(0), 0.0, 1.0)
    zbar = z_k.mean(0)
    zvar = ((K - 1) / K) * np.sum((z_k - zbar[None]) ** 2, axis=0)
    se = np.sqrt(zvar)
    t_crit = _t_dist.ppf(1 - alpha / 2, K - 1)
    zc = fisher_atanh_transform(mean)
    lo = inverse_fisher_atanh(zc - t_crit * se)
    hi = inverse_fisher_atanh(zc + t_crit * se)
    lo = np.minimum(lo, mean)
    hi = np.maximum(hi, mean)
    return mean, lo, hi


# --------------------------------------------------------------------------
# multitaper MSC  (:619-839)
# --------------------------------------------------------------------------
def multitaper_msc(eeg, emg, sampling_freq, nw=3, window_length_sec=1.0, overlap_frac=0.5,
                   taper_eigenvalue_threshold=0.90, use_jackknife=True, jackknife_alpha=0.05,
                   apply_independence_threshold=True, apply_bonferroni_correction=False,
                   significance_level=0.05, window_mask=None):
    """Per-window multitaper MSC with the reference's output dict layout."""
    eeg = np.asarray(eeg, dtype=np.float64)
    emg = np.asarray(emg, dtype=np.float64)
    n, ne = eeg.shape
    n2, nm = emg.shape
    if n != n2:
        raise ValueError("EEG and EMG must have same number of samples.")
    N, hop = window_params(sampling_freq, window_length_sec, overlap_frac)
    tapers, _ = dpss_tapers(N, nw, taper_eigenvalue_threshold)
    K = len(tapers)
    freqs = np.fft.rfftfreq(N, d=1 / sampling_freq)
    F = len(freqs)
    W = n_windows_msc(n, N, hop)
    starts = np.arange(W) * hop
    time_centers = (starts + N / 2) / sampling_freq
    if window_mask is not None:
        window_mask = np.asarray(window_mask, dtype=bool)
        if window_mask.shape != (W,):
            raise ValueError(f"window_mask must have shape ({W},), got {window_mask.shape}")
        active = np.flatnonzero(window_mask)
    else:
        active = np.arange(W)

    coh = np.zeros((W, F, ne, nm), dtype=np.float32)
    lo = np.zeros_like(coh) if use_jackknife else None
    hi = np.zeros_like(coh) if use_jackknife else None
    for w in active:
        X = segment_spectra(eeg, starts[w:w + 1], tapers)[0]        # (K,F,ne)
        Y = segment_spectra(emg, starts[w:w + 1], tapers)[0]
        if use_jackknife:
            m, l, h = jackknife_from_spectra(X, Y, jackknife_alpha)
            coh[w], lo[w], hi[w] = m, l, h
        else:
            coh[w] = msc_from_spectra(X, Y)[0]
    out = {"coherence_raw": coh, "time_centers": time_centers, "freqs": freqs,
           "metadata": {"K_tapers": K, "n_windows": W, "n_active_windows": int(len(active))}}
    if use_jackknife:
        out["coherence_ci_lower"] = lo
        out["coherence_ci_upper"] = hi
    if apply_independence_threshold:
        ncomp = ne * nm if apply_bonferroni_correction else None
        sig = np.zeros(coh.shape, dtype=bool)
        for w in active:
            sig[w], _ = threshold_filtering(coh[w], K, significance_level, ncomp,
                                            apply_bonferroni_correction)
        out["coherence_significant"] = sig
        out["metadata"]["IT_unadjusted"] = independence_threshold(K, significance_level)
    return out


def max_over_emg(cmc, lo=None, hi=None, channel_ax=3):
    """Joint EMG-argmax reduction, :1132-1171."""
    idx = np.argmax(cmc, axis=channel_ax)
    take = lambda a: np.take_along_axis(a, idx[..., None], axis=channel_ax).squeeze(channel_ax)
    if lo is None or hi is None:
        return take(cmc)
    return take(cmc), take(lo), take(hi)


# --------------------------------------------------------------------------
# Welch MSC == scipy.signal.coherence, all pairs  (preprocessing.py:1228-1230)
# --------------------------------------------------------------------------
def welch_segments(n_samples: int, nperseg: int, noverlap: int | None = None):
    if noverlap is None:
        noverlap = nperseg // 2
    step = nperseg - noverlap
    return np.arange(0, n_samples - nperseg + 1, step)


def welch_msc(x, y, nperseg: int, noverlap: int | None = None, window="hann",
              detrend: bool = True, bin_lo: int = 0, bin_hi: int | None = None):
    """All-pairs Welch MSC. x (n, Ne), y (n, Nm) -> (F, Ne, Nm).

    scipy.signal.coherence = |Pxy|^2 / (Pxx Pyy) with Welch estimates: periodic
    hann window, 50 % overlap, per-segment constant detrend before windowing,
    mean over segments.  Window power / fs / one-sided factors cancel.
    """
    x = np.asarray(x, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64)
    starts = welch_segments(x.shape[0], nperseg, noverlap)
    win = signal.get_window(window, nperseg)[None, :]
    d = 1 if detrend else 0
    X = segment_spectra(x, starts, win, d, bin_lo, bin_hi)[:, 0]
    Y = segment_spectra(y, starts, win, d, bin_lo, bin_hi)[:, 0]
    sxx = np.sum(np.abs(X) ** 2, axis=0)
    syy = np.sum(np.abs(Y) ** 2, axis=0)
    sxy = np.einsum("lfi,lfj->fij", np.conj(X), Y)
    with np.errstate(divide="ignore", invalid="ignore"):
        return np.abs(sxy) ** 2 / (sxx[:, :, None] * syy[:, None, :])


# --------------------------------------------------------------------------
# multitaper PSD  (:385-454)
# --------------------------------------------------------------------------
def multitaper_psd(x, sampling_freq, nw=3, window_length_sec=1.0, overlap_frac=0.5,
                   apply_log_scale=True):
    """(W', F, n_ch) spectrogram.  ``signal.periodogram(window=None)`` on the
    already-tapered window (:419) means: boxcar window, detrend='constant'
    applied to the TAPERED data, density scaling 1/(fs N), one-sided doubling
    of every bin except DC and Nyquist; tapers are NOT re-normalised here
    (scipy's dpss rows already have unit L2 norm)."""
    x = np.asarray(x, dtype=np.float64)
    if x.ndim == 1:
        x = x[:, None]
    n = x.shape[0]
    N, hop = window_params(sampling_freq, window_length_sec, overlap_frac)
    k = int(2 * nw - 1)
    tapers = signal.windows.dpss(M=N, NW=nw, Kmax=k)
    starts = window_starts_psd(n, N, hop)
    time_centers = (starts + N / 2) / sampling_freq
    freqs = np.fft.rfftfreq(N, d=1 / sampling_freq)
    S = segment_spectra(x, starts, tapers, detrend=2)            # (W,K,F,C)
    p = np.abs(S) ** 2 / (sampling_freq * N)
    if N % 2 == 0:
        p[:, :, 1:-1] *= 2
    else:
        p[:, :, 1:] *= 2
    spec = p.mean(axis=1)
    if apply_log_scale:
        spec = np.log10(np.abs(spec) + 1e-10)
    return spec, time_centers, freqs
