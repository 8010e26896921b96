"""Drop-in for the coherence part of the reference's ``src/pipeline/signal_features.py``.

Same function names, keyword arguments, defaults, return shapes/dtypes and error behaviour as the
reference (cited per function); the arithmetic runs in the sm_100a kernels of libcmc_b200.so.
Inputs may be numpy arrays (host) or CUDA torch tensors; numpy in -> numpy out.  There is no CPU
path: without a CUDA device or the built library every compute function raises.

New (not in the reference): ``welch_magnitude_squared_coherence`` / ``pooled_coherence`` - the
all-pairs pooled estimator that feeds the surrogate null (``data_surrogation.py``).
"""
from __future__ import annotations

import functools

from pathlib import Path
from typing import Literal

import numpy as np
import torch
from scipy import signal
from scipy.stats import beta, t as t_dist

from . import kernels as K
from .channel_layout import EEG_CHANNEL_IND_DICT
from . import file_management as filemgmt

# signal_features.py:17-26
FREQUENCY_BANDS = {
    'delta': (0.5, 4),
    'theta': (4, 8),
    'alpha': (8, 12),
    'beta': (13, 30),
    'gamma': (30, 100),
}

from . import spectrogram_aggregation as _agg
from .spectrogram_aggregation import aggregate_psd_spectrogram  # noqa: E402,F401  (signal_features.py:1374-1502)


def aggregate_spectrogram_over_frequency_band(spectrograms, freqs, behaviour='mean', frequency_bands=None,
                                              log_transform=False, log_epsilon=1e-10, frequency_axis=1,
                                              pre_aggregate_axis=None, lower_array=None, upper_array=None,
                                              **kwargs):
    """signal_features.py:1174-1371; see ``spectrogram_aggregation`` for the fidelity note on band selection."""
    return _agg.aggregate_spectrogram_over_frequency_band(
        spectrograms, freqs, behaviour, frequency_bands, log_transform, log_epsilon, frequency_axis,
        pre_aggregate_axis, lower_array, upper_array, default_bands=FREQUENCY_BANDS, **kwargs)


# budget (bytes) for the per-call (windows, F, Ne, Nm) device tensors; longer recordings are
# processed in window chunks and streamed to the host arrays
DEVICE_CHUNK_BYTES = 8 << 30


# ----------------------------------------------------------------------------- helpers
def check_2d_numpy_array(input_array, axis: Literal[0, 1] | None = None):
    """signal_features.py:29-37 (AttributeError when a 2-D array comes without axis)."""
    if len(input_array.shape) == 1:
        input_array = input_array[:, np.newaxis]
        if axis is None:
            axis = 0
    else:
        if axis is None:
            raise AttributeError("For 2D signal arrays, axis needs to be defined!")
    return input_array, axis


def _normalize_to_time_first(array, axis: Literal[0, 1]):
    """signal_features.py:1103-1129."""
    if array.ndim != 2:
        raise ValueError(f"Input array must be 2D. Got shape {array.shape}")
    if axis == 0:
        return array
    elif axis == 1:
        return array.T
    else:
        raise ValueError(f"axis must be 0 or 1. Got {axis}")


def _device():
    if not torch.cuda.is_available():
        raise RuntimeError("multimodal_biosignal_analysis_b200 needs a CUDA device (no CPU fallback)")
    return torch.device("cuda", torch.cuda.current_device())


def _to_device_f32(a) -> torch.Tensor:
    """(n_samples, n_ch) float32 CUDA tensor with unit channel stride."""
    if isinstance(a, torch.Tensor):
        # pinned host tensors copy asynchronously, so consecutive uploads queue back to back
        t = a.to(device=_device(), dtype=torch.float32, non_blocking=True)
    else:
        h = np.ascontiguousarray(a, dtype=np.float32)
        t = torch.from_numpy(h).to(_device(), non_blocking=True)
    if t.stride(-1) != 1:
        t = t.contiguous()
    return t


def _is_host(*arrays) -> bool:
    return not any(isinstance(a, torch.Tensor) and a.is_cuda for a in arrays)


def _out(t: torch.Tensor, host: bool):
    return t.cpu().numpy() if host else t


def resample_data(data, original_sampling_freq, new_sampling_freq, axis: Literal[0, 1] = None):
    """Linear-interpolation resampling helper, signal_features.py:40-56 (host utility)."""
    from scipy.interpolate import interp1d
    input_array, axis = check_2d_numpy_array(data, axis=axis)
    n_timesteps = input_array.shape[axis]
    duration = n_timesteps / original_sampling_freq
    new_n = int(round(duration * new_sampling_freq))
    f = interp1d(np.linspace(0, duration, n_timesteps), input_array, axis=axis, kind='linear',
                 fill_value='extrapolate')
    return f(np.linspace(0, duration, new_n))


def mirror_eeg_channel_list(channels: list[str], input_is_left: bool = True) -> list[str]:
    """Left/right mirror of 10-20 channel names, signal_features.py:59-76."""
    out = []
    for ch in channels:
        if ch[-1] == 'z':
            out.append(ch)
            continue
        if ch[-2:].isnumeric():
            idx, area = int(ch[-2:]), ch[:-2]
        elif ch[-1].isnumeric():
            idx, area = int(ch[-1]), ch[:-1]
        else:
            raise ValueError("Unrecognizable EEG channel name: ", ch)
        out.append(f"{area}{idx + (1 if input_is_left else -1)}")
    return out


# ----------------------------------------------------------------------------- PSD
def multitaper_psd(input_array, sampling_freq: float, nw: float = 3, window_length_sec: float = 1.0,
                   overlap_frac: float = 0.5, axis: Literal[0, 1] = None, apply_log_scale: bool = True,
                   psd_save_dir: str | Path | None = None, psd_file_suffix: str = "", plot_result: bool = False,
                   **plot_kwargs):
    """Sliding-window multitaper PSD, signal_features.py:80-454 (body :385-454): windows at
    ``np.arange(0, n - N, hop)`` (:398 - one window fewer than the MSC grid when (n - N) % hop == 0),
    un-renormalised DPSS tapers, ``signal.periodogram(window=None)`` conventions (mean of the TAPERED
    window removed, density scaling 1 / (fs N), one-sided doubling), mean over tapers, optional
    ``log10(|.| + 1e-10)``.  Returns (spectrograms (W, F, n_ch), time_centers, freqs).  ``plot_result`` needs the
    reference's plotting module and is not supported here."""
    host = _is_host(input_array)
    input_array, axis = check_2d_numpy_array(input_array, axis=axis)
    n_samples = input_array.shape[axis]
    window_samples = int(window_length_sec * sampling_freq)
    hop_samples = int(window_samples * (1 - overlap_frac))
    k = int(2 * nw - 1)
    tapers = _dpss_raw(window_samples, nw, k)
    window_starts = np.arange(0, n_samples - window_samples, hop_samples)
    time_centers = (window_starts + window_samples / 2) / sampling_freq
    freqs = np.fft.rfftfreq(window_samples, d=1 / sampling_freq)
    if axis == 1:
        input_array = input_array.T
    dev = _device()
    x = _to_device_f32(input_array)
    starts_d = torch.from_numpy(window_starts.astype(np.int64)).to(dev)
    td = torch.from_numpy(np.ascontiguousarray(tapers, dtype=np.float32)).to(dev)
    spec = K.fft_segments(x, starts_d, td, K.DETREND_POST_TAPER)
    psd = K.psd_from_spectra(spec, 1.0 / (sampling_freq * window_samples), True, 0, window_samples, apply_log_scale)
    spectrograms = _out(psd, host)
    if host:
        spectrograms = spectrograms.astype(np.float64)
    if psd_save_dir is not None:
        save_spectrograms(spectrograms, time_centers, freqs, "PSD", save_dir=psd_save_dir,
                          identifier_suffix=psd_file_suffix)
    if plot_result:
        raise NotImplementedError("plot_result needs the reference's visualizations module (out of scope)")
    return spectrograms, time_centers, freqs


def welch_psd(input_array, sampling_freq: float, nperseg: int, noverlap: int | None = None, window: str = "hann",
              axis: Literal[0, 1] = 0):
    """``scipy.signal.welch(x, fs, nperseg=...)`` defaults (hann, 50 % overlap, constant detrend, density,
    one-sided, mean over segments) for every channel: returns (freqs, psd (F, n_ch))."""
    host = _is_host(input_array)
    input_array, axis = check_2d_numpy_array(input_array, axis=axis)
    if axis == 1:
        input_array = input_array.T
    n = input_array.shape[0]
    nperseg = int(min(nperseg, n))                      # scipy clips nperseg to the input length
    if noverlap is None:
        noverlap = nperseg // 2
    starts = np.arange(0, n - nperseg + 1, nperseg - noverlap, dtype=np.int64)
    win = signal.get_window(window, nperseg)
    dev = _device()
    spec = K.fft_segments(_to_device_f32(input_array), torch.from_numpy(starts).to(dev),
                          torch.from_numpy(win.astype(np.float32)[None]).to(dev), K.DETREND_CONSTANT)
    L, _, F, C = spec.shape
    psd = K.psd_from_spectra(spec.view(1, L, F, C), 1.0 / (sampling_freq * float(np.sum(win ** 2))), True, 0, nperseg,
                             False)[0]
    freqs = np.fft.rfftfreq(nperseg, d=1 / sampling_freq)
    return freqs, (_out(psd, host).astype(np.float64) if host else psd)


def compute_spectral_snr(input_array, sampling_freq: int, target_freq: float = 21.5, freq_window: float = 8.5,
                         target_band_ratio: float = .5, axis: Literal[0, 1] = 0, return_psd: bool = False):
    """Spectral SNR (dB) at a target frequency from a Welch PSD with 4-second segments,
    signal_features.py:2069-2130: mean PSD in the narrow target band over mean PSD in the noise band
    (means over bins AND channels, like the reference's boolean indexing of the (F, n_ch) array)."""
    input_array, axis = check_2d_numpy_array(input_array, axis=axis)
    freqs, psd = welch_psd(input_array, sampling_freq, nperseg=sampling_freq * 4, axis=axis)
    if isinstance(psd, torch.Tensor):
        psd = psd.cpu().numpy().astype(np.float64)
    if axis == 1:
        psd = psd.T                                     # scipy keeps the time axis in place
    fax = freqs
    target_freq_window = freq_window * target_band_ratio
    target_band = (fax < target_freq + target_freq_window) & (fax > target_freq - target_freq_window)
    noise_band = (fax >= target_freq - freq_window) & (fax <= target_freq + freq_window)
    if axis == 1:
        snr_linear = np.mean(psd[:, target_band]) / np.mean(psd[:, noise_band])
    else:
        snr_linear = np.mean(psd[target_band]) / np.mean(psd[noise_band])
    snr_db = 10 * np.log10(snr_linear)
    return snr_db if not return_psd else (snr_db, freqs, psd)


# ----------------------------------------------------------------------------- scalar statistics
def fisher_atanh_transform(coherence, eps: float = 1e-10):
    """signal_features.py:459-462."""
    c = np.clip(coherence, eps, 1 - eps)
    return 0.5 * np.log((1 + c) / (1 - c))


def inverse_fisher_atanh(z):
    """signal_features.py:465-467 (tanh(z)**2 - not the algebraic inverse)."""
    return np.tanh(z) ** 2


def compute_cmc_independence_threshold(K: int, alpha: float = 0.05) -> float:
    """(1 - alpha) quantile of Beta(K-2, K-2), signal_features.py:470-481."""
    a = b = K - 2
    return beta.ppf(1 - alpha, a, b)


def apply_threshold_filtering(coherence_values, K: int, alpha: float = 0.05, n_comparisons: int = None,
                              apply_bonferroni: bool = False) -> tuple:
    """signal_features.py:581-604."""
    if apply_bonferroni and n_comparisons is not None:
        alpha_adjusted = alpha / n_comparisons
        if alpha_adjusted < 1e-10:
            alpha_adjusted = 1e-10
    else:
        alpha_adjusted = alpha
    IT = compute_cmc_independence_threshold(K, alpha=alpha_adjusted)
    return coherence_values > IT, IT


@functools.lru_cache(maxsize=32)
def _dpss_raw(window_samples: int, nw: float, k: int):
    """Un-renormalised DPSS rows as multitaper_psd uses them (signal_features.py:395), cached like _dpss."""
    out = signal.windows.dpss(M=window_samples, NW=nw, Kmax=k)
    out.setflags(write=False)
    return out


@functools.lru_cache(maxsize=32)
def _dpss(window_samples: int, nw: float, eig_threshold: float):
    """Tapers kept by eigenvalue and L2-normalised, signal_features.py:669-678.  Host side; scipy needs ~16 ms for
    N = 4096 - more than the GPU needs for a ten-minute recording - so the (read-only) table is cached."""
    k = int(2 * nw - 1)
    tapers, eigs = signal.windows.dpss(M=window_samples, NW=nw, Kmax=k, return_ratios=True)
    kept = tapers[eigs > eig_threshold]
    out = np.stack([t / np.sqrt(np.sum(t ** 2)) for t in kept]) if len(kept) else kept
    out.setflags(write=False)
    return out


# ----------------------------------------------------------------------------- multitaper MSC
def multitaper_magnitude_squared_coherence(
        eeg_array,
        emg_array,
        sampling_freq: float,
        nw: float = 3,
        window_length_sec: float = 1.0,
        overlap_frac: float = 0.5,
        eeg_axis: Literal[0, 1] = 0,
        emg_axis: Literal[0, 1] = 0,
        taper_eigenvalue_threshold: float = 0.90,
        use_jackknife: bool = True,
        jackknife_alpha: float = 0.05,
        apply_independence_threshold: bool = True,
        apply_bonferroni_correction: bool = False,
        significance_level: float = 0.05,
        window_mask=None,
        verbose: bool = False,
        *,
        freq_band: tuple[float, float] | None = None,
        reduce_emg: bool = False,
        zero_nonsignificant: bool = False,
) -> dict:
    """Sliding-window multitaper MSC, signal_features.py:619-839.

    Returns the reference's dict: ``coherence_raw`` (W, F, Ne, Nm) float32 (jackknife mean when
    ``use_jackknife``), ``time_centers``, ``freqs``, optional ``coherence_ci_lower/upper``,
    ``coherence_significant`` (bool) and ``metadata``.  Masked-out windows stay zero but keep
    their time centre (:727-733).

    Keyword-only extensions (defaults reproduce the reference): ``freq_band=(lo, hi)`` keeps only
    bins with lo <= f <= hi; ``reduce_emg=True`` fuses ``max_cmc_spectrograms_over_channels`` so
    the outputs have shape (W, F, Ne) and the 4-D tensor never exists (adds ``emg_argmax``).
    """
    host = _is_host(eeg_array, emg_array)
    eeg_array = _normalize_to_time_first(eeg_array, axis=eeg_axis)
    emg_array = _normalize_to_time_first(emg_array, axis=emg_axis)
    n_samples_eeg, n_eeg_channels = eeg_array.shape
    n_samples_emg, n_emg_channels = emg_array.shape
    if n_samples_eeg != n_samples_emg:
        raise ValueError(
            f"EEG and EMG must have same number of samples. "
            f"Got EEG: {n_samples_eeg}, EMG: {n_samples_emg}"
        )
    n_samples = n_samples_eeg

    window_samples = int(window_length_sec * sampling_freq)
    hop_samples = int(window_samples * (1 - overlap_frac))
    tapers = _dpss(window_samples, nw, taper_eigenvalue_threshold)
    n_tapers = len(tapers)
    freqs = np.fft.rfftfreq(window_samples, d=1 / sampling_freq)
    n_windows = (n_samples - window_samples) // hop_samples + 1

    if window_mask is not None:
        window_mask = np.asarray(window_mask.cpu() if isinstance(window_mask, torch.Tensor) else window_mask,
                                 dtype=bool)
        if window_mask.shape != (n_windows,):
            raise ValueError(f"window_mask must have shape ({n_windows},), got {window_mask.shape}")
        n_active = int(window_mask.sum())
    else:
        n_active = n_windows
    if verbose:
        print(f"Using {n_tapers} high-quality tapers (λ > {taper_eigenvalue_threshold})")
        print(f"Computing MSC: {n_eeg_channels} EEG × {n_emg_channels} EMG channels")

    bin_lo, bin_hi = 0, len(freqs) - 1
    if freq_band is not None:
        sel = np.flatnonzero((freqs >= freq_band[0]) & (freqs <= freq_band[1]))
        if len(sel) == 0:
            raise ValueError(f"freq_band {freq_band} selects no frequency bin")
        bin_lo, bin_hi = int(sel[0]), int(sel[-1])
        freqs = freqs[bin_lo:bin_hi + 1]
    n_freqs = bin_hi - bin_lo + 1

    it_threshold = None
    if apply_independence_threshold:
        n_comp = n_eeg_channels * n_emg_channels if apply_bonferroni_correction else None
        _, it_threshold = apply_threshold_filtering(np.zeros(1), K=n_tapers, alpha=significance_level,
                                                    n_comparisons=n_comp,
                                                    apply_bonferroni=apply_bonferroni_correction)
    t_crit = float(t_dist.ppf(1 - (jackknife_alpha / 2), n_tapers - 1)) if use_jackknife else 0.0

    dev = _device()
    eeg_d = _to_device_f32(eeg_array)
    emg_d = _to_device_f32(emg_array)
    tapers_d = torch.from_numpy(np.ascontiguousarray(tapers, dtype=np.float32)).to(dev)
    starts_all = np.arange(n_windows, dtype=np.int64) * hop_samples
    time_centers = (starts_all + window_samples / 2) / sampling_freq

    out_shape = (n_windows, n_freqs, n_eeg_channels) if reduce_emg else \
        (n_windows, n_freqs, n_eeg_channels, n_emg_channels)
    n_arrays = 1 + (2 if use_jackknife else 0)
    per_window = int(np.prod(out_shape[1:])) * (4 * n_arrays + 1) + \
        n_tapers * n_freqs * (n_eeg_channels + n_emg_channels) * 8
    chunk = max(1, min(n_windows, DEVICE_CHUNK_BYTES // max(per_window, 1), 65535))
    single = chunk >= n_windows

    def alloc(dtype):
        if single:
            return None
        if host:
            return np.zeros(out_shape, dtype=dtype)
        return torch.zeros(out_shape, dtype=getattr(torch, np.dtype(dtype).name) if dtype != bool else torch.bool,
                           device=dev)

    res = {"coherence_raw": alloc(np.float32)}
    if use_jackknife:
        res["coherence_ci_lower"] = alloc(np.float32)
        res["coherence_ci_upper"] = alloc(np.float32)
    if apply_independence_threshold and not reduce_emg:
        res["coherence_significant"] = alloc(bool)
    if reduce_emg:
        res["emg_argmax"] = alloc(np.int32)

    for w0 in range(0, n_windows, chunk):
        w1 = min(n_windows, w0 + chunk)
        starts_d = torch.from_numpy(starts_all[w0:w1]).to(dev)
        mask_d = None
        if window_mask is not None:
            mask_d = torch.from_numpy(window_mask[w0:w1].astype(np.uint8)).to(dev)
        X = K.fft_segments(eeg_d, starts_d, tapers_d, K.DETREND_NONE, bin_lo, bin_hi)
        Y = K.fft_segments(emg_d, starts_d, tapers_d, K.DETREND_NONE, bin_lo, bin_hi)
        if reduce_emg:
            coh, lo, hi, arg = K.msc_windows_maxemg(X, Y, mask_d, use_jackknife, t_crit, it_threshold,
                                                    zero_nonsignificant, True)
            parts = {"coherence_raw": coh, "coherence_ci_lower": lo, "coherence_ci_upper": hi, "emg_argmax": arg}
        else:
            coh, lo, hi, sig = K.msc_windows(X, Y, mask_d, use_jackknife, t_crit, it_threshold)
            parts = {"coherence_raw": coh, "coherence_ci_lower": lo, "coherence_ci_upper": hi,
                     "coherence_significant": None if sig is None else sig.bool()}
        for key, val in parts.items():
            if key not in res or val is None:
                continue
            if single:
                res[key] = _out(val, host)
            elif host:
                res[key][w0:w1] = val.cpu().numpy()
            else:
                res[key][w0:w1] = val

    result = {
        "coherence_raw": res["coherence_raw"],
        "time_centers": time_centers,
        "freqs": freqs,
        "metadata": {
            "K_tapers": n_tapers,
            "n_windows": n_windows,
            "n_active_windows": n_active,
            "window_length_sec": window_length_sec,
            "overlap_frac": overlap_frac,
            "use_jackknife": use_jackknife,
            "apply_independence_threshold": apply_independence_threshold,
            "apply_bonferroni_correction": apply_bonferroni_correction,
            "significance_level": significance_level,
        },
    }
    if use_jackknife:
        result["coherence_ci_lower"] = res["coherence_ci_lower"]
        result["coherence_ci_upper"] = res["coherence_ci_upper"]
    if reduce_emg:
        result["emg_argmax"] = res["emg_argmax"]
    if apply_independence_threshold:
        IT_unadjusted = compute_cmc_independence_threshold(n_tapers, alpha=significance_level)
        result["metadata"]["IT_unadjusted"] = float(IT_unadjusted)
        if apply_bonferroni_correction:
            n_comp = n_eeg_channels * n_emg_channels
            result["metadata"]["IT_bonferroni"] = float(
                compute_cmc_independence_threshold(n_tapers, alpha=significance_level / n_comp))
            result["metadata"]["n_comparisons"] = n_comp
        if not reduce_emg:
            result["coherence_significant"] = res["coherence_significant"]
            result["metadata"]["n_significant"] = int(res["coherence_significant"].sum())
    if verbose:
        print("\n✓ Done!")
    return result


def jackknife_coherence_and_ci(tapers_filtered, eeg_window, emg_window, sampling_freq: float,
                               window_samples: int, jackknife_alpha: float = 0.05) -> tuple:
    """Leave-one-taper-out mean and Student-t CI of one window, signal_features.py:484-578.
    Returns (coherence_mean, ci_lower, ci_upper), each (F, Ne, Nm) float32."""
    host = _is_host(eeg_window, emg_window)
    tapers = np.ascontiguousarray(np.stack([np.asarray(t) for t in tapers_filtered]), dtype=np.float32)
    n_tapers = len(tapers)
    if tapers.shape[1] != window_samples or eeg_window.shape[0] != window_samples:
        raise ValueError("taper / window length mismatch")
    dev = _device()
    starts = torch.zeros(1, dtype=torch.int64, device=dev)
    td = torch.from_numpy(tapers).to(dev)
    X = K.fft_segments(_to_device_f32(eeg_window), starts, td, K.DETREND_NONE)
    Y = K.fft_segments(_to_device_f32(emg_window), starts, td, K.DETREND_NONE)
    t_crit = float(t_dist.ppf(1 - (jackknife_alpha / 2), n_tapers - 1))
    coh, lo, hi, _ = K.msc_windows(X, Y, None, True, t_crit, None)
    return _out(coh[0], host), _out(lo[0], host), _out(hi[0], host)


def max_cmc_spectrograms_over_channels(cmc_array, cmc_array_lower_ci=None, cmc_array_upper_ci=None,
                                       channel_ax: int = 3, verbose: bool = True):
    """Joint EMG-argmax gather, signal_features.py:1132-1171 (index/gather glue on already
    materialised arrays; the fused device path is ``reduce_emg=True`` above)."""
    if verbose:
        print("Maxing CMC values over EMG channels (aligned)...")
    if isinstance(cmc_array, torch.Tensor):
        # first index on ties like np.argmax: max value, then smallest index among equals
        mx = cmc_array.amax(dim=channel_ax, keepdim=True)
        n = cmc_array.shape[channel_ax]
        shape = [1] * cmc_array.dim()
        shape[channel_ax] = n
        ar = torch.arange(n, device=cmc_array.device).view(shape)
        idx = torch.where(cmc_array == mx, ar, n).amin(dim=channel_ax, keepdim=True)
        take = lambda a: torch.take_along_dim(a, idx, dim=channel_ax).squeeze(channel_ax)
    else:
        idx = np.argmax(cmc_array, axis=channel_ax)[..., np.newaxis]
        take = lambda a: np.take_along_axis(a, idx, axis=channel_ax).squeeze(axis=channel_ax)
    maxed = take(cmc_array)
    if cmc_array_lower_ci is None or cmc_array_upper_ci is None:
        return maxed
    return maxed, take(cmc_array_lower_ci), take(cmc_array_upper_ci)


def _build_task_window_mask(time_centers_sec, log_frame, pre_buffer_sec: float, post_buffer_sec: float):
    """signal_features.py:842-895.  The trial-log readers come from ``experiment_log`` through the
    ``data_integration`` namespace below (test doubles may be patched onto it)."""
    import pandas as pd
    measurement_start, _ = data_integration.get_qtc_measurement_start_end(log_frame)
    measurement_start_aware = pd.Timestamp(measurement_start)
    if measurement_start_aware.tzinfo is None:      # data_analysis.make_timezone_aware: naive -> UTC
        measurement_start_aware = measurement_start_aware.tz_localize('utc')
    trial_start_ends = data_integration.get_all_task_start_ends(log_frame, output_type='list')
    mask = np.zeros(len(time_centers_sec), dtype=bool)
    for trial_start, trial_end in trial_start_ends:
        t0 = (trial_start - measurement_start_aware).total_seconds() - pre_buffer_sec
        t1 = (trial_end - measurement_start_aware).total_seconds() + post_buffer_sec
        mask |= (time_centers_sec >= t0) & (time_centers_sec <= t1)
    n_active = int(mask.sum())
    print(f"Task window mask: {n_active}/{len(mask)} windows selected "
          f"({100 * n_active / len(mask):.1f}%) across {len(trial_start_ends)} trials "
          f"[±{pre_buffer_sec}s / +{post_buffer_sec}s buffers]")
    return mask


class _DataIntegrationShim:
    """Namespace with the two trial-log readers ``_build_task_window_mask`` needs (restated in
    ``experiment_log.py`` after ``src/pipeline/data_integration.py:717,766``); tests and callers may replace them
    (``monkeypatch.setattr(features.data_integration, ...)`` as the reference's tests do)."""
    from . import experiment_log as _log
    get_all_task_start_ends = staticmethod(_log.get_all_task_start_ends)
    get_qtc_measurement_start_end = staticmethod(_log.get_qtc_measurement_start_end)


data_integration = _DataIntegrationShim()


def compute_task_wise_aggregated_cmc(
        eeg_array,
        emg_array,
        sampling_freq: int,
        muscle_group: str,
        log_frame=None,
        eeg_channel_subset: list[str] | None = None,
        window_size_sec: float = 2.0,
        window_overlap_ratio: float = 0.5,
        enforce_independence_threshold: bool = False,
        independence_threshold_alpha: float = 0.2,
        use_jackknife: bool = True,
        jackknife_alpha: float = 0.05,
        save_dir: str | Path | None = None,
        pre_trial_computation_buffer_sec: float = 3.0,
        post_trial_computation_buffer_sec: float = 3.0,
) -> tuple:
    """Channel-aggregated CMC, signal_features.py:898-1026.  The EMG-argmax reduction
    (:1132-1171), the significance zeroing (:979-983) and the CI gather run fused on the device;
    the (W, F, Ne, Nm) tensors of the reference are never materialised.  The CI-ordering asserts
    (:986-990) hold by construction (kernel clamps lower <= mean <= upper)."""
    if eeg_channel_subset:
        eeg_channel_subset_inds = [EEG_CHANNEL_IND_DICT[ch] for ch in eeg_channel_subset]
        print(f"Reducing EEG to {len(eeg_channel_subset)} channels: {eeg_channel_subset}")
        eeg_array = eeg_array[:, eeg_channel_subset_inds]
    n_samples_eeg, n_eeg_channels = eeg_array.shape
    n_samples_emg, n_emg_channels = emg_array.shape
    if n_samples_eeg != n_samples_emg:
        raise ValueError(
            f"EEG and EMG must have same number of samples. "
            f"Got EEG: {n_samples_eeg}, EMG: {n_samples_emg}"
        )
    if log_frame is not None:
        window_samples = int(window_size_sec * sampling_freq)
        hop_samples = int(window_samples * (1 - window_overlap_ratio))
        if hop_samples <= 0:
            raise ValueError("window_overlap_ratio too high: hop_samples becomes <= 0")
        n_windows = (n_samples_eeg - window_samples) // hop_samples + 1
        window_starts = np.arange(n_windows) * hop_samples
        time_centers_preview = (window_starts + window_samples / 2) / sampling_freq
        window_mask = _build_task_window_mask(
            time_centers_sec=time_centers_preview, log_frame=log_frame,
            pre_buffer_sec=pre_trial_computation_buffer_sec,
            post_buffer_sec=post_trial_computation_buffer_sec)
    else:
        window_mask = None

    output_dict = multitaper_magnitude_squared_coherence(
        eeg_array, emg_array,
        sampling_freq=sampling_freq,
        window_length_sec=window_size_sec,
        overlap_frac=window_overlap_ratio,
        significance_level=independence_threshold_alpha,
        apply_independence_threshold=enforce_independence_threshold,
        use_jackknife=use_jackknife,
        jackknife_alpha=jackknife_alpha,
        window_mask=window_mask,
        verbose=True,
        reduce_emg=True,
        zero_nonsignificant=enforce_independence_threshold,
    )
    time_centers = output_dict['time_centers']
    freqs = output_dict['freqs']
    values = output_dict['coherence_raw']
    if use_jackknife:
        values_lower = output_dict['coherence_ci_lower']
        values_upper = output_dict['coherence_ci_upper']
    if np.ndim(values) == 4:
        # a caller-substituted MSC function returned the un-reduced (W, F, Ne, Nm) tensors:
        # reduce them exactly like the reference does (:975-1002)
        if enforce_independence_threshold:
            values = np.where(output_dict['coherence_significant'], values, 0.0)
        if use_jackknife:
            values, values_lower, values_upper = max_cmc_spectrograms_over_channels(
                values, values_lower, values_upper, channel_ax=3, verbose=True)
        else:
            values = max_cmc_spectrograms_over_channels(values, channel_ax=3, verbose=True)

    if save_dir is not None:
        channel_suffix = (f"Channels_{'_'.join(eeg_channel_subset)}" if eeg_channel_subset else "All_Channels")
        label = f"{muscle_group.capitalize()} CMC{' Trial-wise' if log_frame is not None else ''}"
        save_spectrograms(values, time_centers, freqs, save_dir=save_dir, modality=label,
                          identifier_suffix=channel_suffix)
    if use_jackknife:
        return values, values_lower, values_upper, time_centers, freqs
    return values, time_centers, freqs


# ----------------------------------------------------------------------------- pooled estimator (new)
class PooledCoherence:
    """All-pairs pooled magnitude-squared coherence (Welch segments or windows x tapers) on the
    tensor cores.  Keeps the whitened operands on the device so the surrogate null
    (``data_surrogation.surrogate_null``) costs one contraction per distinct surrogate."""

    def __init__(self, csd: K.PooledCsd, freqs: np.ndarray, group: int, host: bool):
        self._csd, self.freqs, self.group, self._host = csd, freqs, group, host

    @property
    def device_result(self) -> K.PooledCsd:
        return self._csd

    @property
    def coherence(self):
        return _out(self._csd.coh, self._host)

    @property
    def sxx(self):
        return _out(self._csd.sxx, self._host)

    @property
    def syy(self):
        return _out(self._csd.syy, self._host)

    @property
    def n_terms(self) -> int:
        return self._csd.dims[0]


def _is_periodic_hann(windows) -> bool:
    """Whether ``windows`` is ONE row equal to scipy's default Welch window (periodic hann, float32): the case the
    tensor-core Welch kernel computes without an FFT."""
    w = np.atleast_2d(np.asarray(windows))
    if w.shape[0] != 1:
        return False
    return bool(np.array_equal(w[0].astype(np.float32), signal.get_window("hann", w.shape[1]).astype(np.float32)))


def pooled_coherence(eeg_array, emg_array, sampling_freq: float, segment_starts, windows,
                     detrend: int = K.DETREND_CONSTANT, freq_band: tuple[float, float] | None = None,
                     eeg_axis: Literal[0, 1] = 0, emg_axis: Literal[0, 1] = 0) -> PooledCoherence:
    """Coherence pooled over ``len(segment_starts) * len(windows)`` spectral estimates.
    windows (K, N): one hann row = Welch; K DPSS rows = multitaper pooled over windows x tapers."""
    host = _is_host(eeg_array, emg_array)
    eeg_array = _normalize_to_time_first(eeg_array, axis=eeg_axis)
    emg_array = _normalize_to_time_first(emg_array, axis=emg_axis)
    if eeg_array.shape[0] != emg_array.shape[0]:
        raise ValueError(
            f"EEG and EMG must have same number of samples. "
            f"Got EEG: {eeg_array.shape[0]}, EMG: {emg_array.shape[0]}")
    windows = np.atleast_2d(np.asarray(windows, dtype=np.float32))
    n_win, N = windows.shape
    freqs = np.fft.rfftfreq(N, d=1 / sampling_freq)
    lo, hi = 0, len(freqs) - 1
    if freq_band is not None:
        sel = np.flatnonzero((freqs >= freq_band[0]) & (freqs <= freq_band[1]))
        if len(sel) == 0:
            raise ValueError(f"freq_band {freq_band} selects no frequency bin")
        lo, hi = int(sel[0]), int(sel[-1])
    dev = _device()
    K.check_segments(segment_starts, N, eeg_array.shape[0])
    starts_d = torch.as_tensor(np.asarray(segment_starts, dtype=np.int64)).to(dev)
    wd = torch.from_numpy(windows).to(dev)
    # K2 contracts the spectra as K1 writes them (no pack pass); its TMA needs an even channel pitch, so odd
    # channel counts get one padding column.  Both modalities land in ONE array (EEG columns, then EMG columns) so
    # that one K1 launch transforms them together.  The operand planes of the surrogate nulls are built on demand.
    eeg_d, emg_d = _to_device_f32(eeg_array), _to_device_f32(emg_array)
    ne, nm = eeg_d.shape[1], emg_d.shape[1]
    ne_p, nm_p = ne + (ne & 1), nm + (nm & 1)
    spec = torch.empty((len(segment_starts), n_win, hi - lo + 1, ne_p + nm_p), dtype=torch.complex64, device=dev)
    K.welch_spectra_pair(eeg_d, emg_d, segment_starts, starts_d, wd, _is_periodic_hann(windows), detrend, lo, hi,
                         spec[..., :ne], spec[..., ne_p:ne_p + nm])
    flat = spec.view(-1, spec.shape[2], spec.shape[3])
    csd = K.csd_msc(flat[:, :, :ne], flat[:, :, ne_p:ne_p + nm])
    return PooledCoherence(csd, freqs[lo:hi + 1], n_win, host)


def welch_magnitude_squared_coherence(eeg_array, emg_array, sampling_freq: float, nperseg: int = 256,
                                      noverlap: int | None = None, window: str = "hann",
                                      detrend: str | bool = "constant",
                                      freq_band: tuple[float, float] | None = None, segment_starts=None,
                                      eeg_axis: Literal[0, 1] = 0, emg_axis: Literal[0, 1] = 0) -> PooledCoherence:
    """All-pairs equivalent of ``scipy.signal.coherence(x, y, fs, nperseg=...)`` as the reference
    uses it (preprocessing.py:1228-1230): periodic hann, 50 % overlap, per-segment constant detrend.
    ``segment_starts`` overrides the regular grid (e.g. segments that must not straddle epochs)."""
    n = eeg_array.shape[eeg_axis]
    if noverlap is None:
        noverlap = nperseg // 2
    if segment_starts is None:
        segment_starts = np.arange(0, n - nperseg + 1, nperseg - noverlap, dtype=np.int64)
    win = signal.get_window(window, nperseg)
    if detrend not in ("constant", False, None):
        raise ValueError("only detrend='constant' or False are supported")
    d = K.DETREND_CONSTANT if detrend == "constant" else K.DETREND_NONE
    return pooled_coherence(eeg_array, emg_array, sampling_freq, segment_starts, win[None], d, freq_band,
                            eeg_axis, emg_axis)


class _RecordingPipeline:
    """Streams recordings of equal shape through the device: upload of item i + 1, compute of item i and download of
    item i - 1 overlap on three CUDA streams, three buffer sets rotate (one uploading, one computing / downloading,
    one in the caller's hands).  ``compute(slot, index)`` is called on the compute stream with ``slot["eeg"]``,
    ``slot["emg"]`` resident (float32 (n, channels)) and returns ``{name: device tensor}`` matching ``out_specs``
    (``{name: (shape, torch dtype)}``); ``slot`` is a dict the callback may keep per-slot scratch in.

    Items are ``(eeg, emg)``, time-first.  Pinned float32 host tensors upload asynchronously as they are; CUDA tensors
    are used in place; anything else (numpy, pageable, float64, ...) is converted into a pinned staging buffer first
    by a few host threads (numpy releases the GIL for the copy) - the extra host pass the reference's callers pay
    when they hand over what ``np.load`` returned."""

    N_SLOTS = 3
    _pool = None
    _pinned_free: dict = {}          # (shape, dtype) -> idle pinned host tensors; cudaHostAlloc of a 126 MB
    # recording costs tens of milliseconds, so the buffers of a finished sweep are kept for the next one

    @classmethod
    def _pinned(cls, shape, dtype):
        free = cls._pinned_free.get((tuple(shape), dtype))
        return free.pop() if free else torch.empty(tuple(shape), dtype=dtype).pin_memory()

    @classmethod
    def _release(cls, t):
        cls._pinned_free.setdefault((tuple(t.shape), t.dtype), []).append(t)

    def __init__(self, n: int, ne: int, nm: int, out_specs: dict, compute, stage_threads: int = 4):
        self.n, self.ne, self.nm, self.compute, self.stage_threads = n, ne, nm, compute, stage_threads
        dev = _device()
        self.slots = []
        for _ in range(self.N_SLOTS):
            self.slots.append(dict(
                eeg=torch.empty((n, ne), dtype=torch.float32, device=dev),
                emg=torch.empty((n, nm), dtype=torch.float32, device=dev),
                out={k: self._pinned(shape, dt) for k, (shape, dt) in out_specs.items()},
                stage=None, h2d=None, comp=None, d2h=None, res=None, src=None))
        self.s_up, self.s_comp, self.s_down = torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.Stream()
        cur = torch.cuda.current_stream()
        for st in (self.s_up, self.s_comp, self.s_down):
            st.wait_stream(cur)

    @classmethod
    def _threads(cls):
        if cls._pool is None:
            import concurrent.futures as cf
            cls._pool = cf.ThreadPoolExecutor(max_workers=8, thread_name_prefix="cmc-stage")
        return cls._pool

    def _host_f32(self, a, slot, key):
        if isinstance(a, torch.Tensor):
            if a.is_cuda:
                return a
            if a.dtype == torch.float32 and a.is_pinned() and a.is_contiguous():
                return a
            a = a.numpy()
        if slot["stage"] is None:
            slot["stage"] = {"eeg": self._pinned((self.n, self.ne), torch.float32),
                             "emg": self._pinned((self.n, self.nm), torch.float32)}
        buf = slot["stage"][key]
        dst, src = buf.numpy(), np.asarray(a)
        nt = self.stage_threads if src.size >= (1 << 20) else 1
        if nt <= 1:
            np.copyto(dst, src, casting="same_kind")
        else:
            edges = np.linspace(0, self.n, nt + 1).astype(np.int64)
            futs = [self._threads().submit(np.copyto, dst[b:e], src[b:e], "same_kind")
                    for b, e in zip(edges[:-1], edges[1:]) if e > b]
            for f in futs:
                f.result()
        return buf

    def _submit(self, item, i):
        slot = self.slots[i % self.N_SLOTS]
        eeg, emg = item
        if tuple(eeg.shape) != (self.n, self.ne) or tuple(emg.shape) != (self.n, self.nm):
            raise ValueError("all recordings of a sweep must have the shape of the first one")
        if slot["comp"] is not None:                              # the slot's previous item has been transformed
            self.s_up.wait_event(slot["comp"])
        if slot["stage"] is not None and slot["h2d"] is not None:
            slot["h2d"].synchronize()                             # staging buffer still being read by the copy engine
        he, hm = self._host_f32(eeg, slot, "eeg"), self._host_f32(emg, slot, "emg")
        slot["src"] = (he, hm)                                    # keep the sources alive until the copy has run
        with torch.cuda.stream(self.s_up):
            slot["eeg"].copy_(he, non_blocking=True)
            slot["emg"].copy_(hm, non_blocking=True)
            slot["h2d"] = torch.cuda.Event()
            slot["h2d"].record()
        with torch.cuda.stream(self.s_comp):
            self.s_comp.wait_event(slot["h2d"])
            if slot["d2h"] is not None:
                self.s_comp.wait_event(slot["d2h"])               # previous results of this slot have left the device
            res = self.compute(slot, i)
            slot["comp"] = torch.cuda.Event()
            slot["comp"].record()
        with torch.cuda.stream(self.s_down):
            self.s_down.wait_event(slot["comp"])
            for k, t in res.items():
                slot["out"][k].copy_(t, non_blocking=True)
                t.record_stream(self.s_down)
            slot["d2h"] = torch.cuda.Event()
            slot["d2h"].record()
        slot["res"] = res

    def _collect(self, i):
        slot = self.slots[i % self.N_SLOTS]
        slot["d2h"].synchronize()
        return {k: t.numpy() for k, t in slot["out"].items()}

    def run(self, first, rest):
        """Generator over ``[first] + rest``: yields the dict of host arrays of every item, in order.  An array
        stays valid while the NEXT item is fetched and is overwritten when the one after that is requested."""
        try:
            self._submit(first, 0)
            i = 0
            for item in rest:
                i += 1
                self._submit(item, i)                             # item i is in flight while item i - 1 is handed out
                yield self._collect(i - 1)
            yield self._collect(i)
        finally:
            # also when the consumer stops early: nothing of this sweep may still be running when its buffers are freed
            for st in (self.s_up, self.s_comp, self.s_down):
                st.synchronize()
            # staging buffers go back to the pool; the OUTPUT buffers do not - the caller may still hold views of them
            for slot in self.slots:
                if slot["stage"] is not None:
                    for t in slot["stage"].values():
                        self._release(t)
                    slot["stage"] = None


def _welch_plan(n: int, sampling_freq: float, nperseg: int, noverlap, window: str, detrend, freq_band, segment_starts):
    """Shared argument handling of the Welch entry points: (segment_starts, window row, detrend mode, lo, hi, freqs)."""
    if noverlap is None:
        noverlap = nperseg // 2
    if segment_starts is None:
        segment_starts = np.arange(0, n - nperseg + 1, nperseg - noverlap, dtype=np.int64)
    if detrend not in ("constant", False, None):
        raise ValueError("only detrend='constant' or False are supported")
    dmode = K.DETREND_CONSTANT if detrend == "constant" else K.DETREND_NONE
    freqs = np.fft.rfftfreq(nperseg, d=1 / sampling_freq)
    lo, hi = 0, len(freqs) - 1
    if freq_band is not None:
        sel = np.flatnonzero((freqs >= freq_band[0]) & (freqs <= freq_band[1]))
        if len(sel) == 0:
            raise ValueError(f"freq_band {freq_band} selects no frequency bin")
        lo, hi = int(sel[0]), int(sel[-1])
    K.check_segments(segment_starts, nperseg, n)
    win = signal.get_window(window, nperseg).astype(np.float32)[None]
    return np.asarray(segment_starts, dtype=np.int64), win, dmode, lo, hi, freqs


def _first_item(recordings):
    it = iter(recordings)
    try:
        first = next(it)
    except StopIteration:
        return None, it
    eeg0, emg0 = first
    if int(emg0.shape[0]) != int(eeg0.shape[0]):
        raise ValueError("EEG and EMG must have same number of samples")
    return first, it


def welch_coherence_sweep(recordings, sampling_freq: float, nperseg: int = 256, noverlap: int | None = None,
                          window: str = "hann", detrend: str | bool = "constant",
                          freq_band: tuple[float, float] | None = None, segment_starts=None):
    """Welch coherence of MANY recordings of equal shape (the subject-condition sweep of the workflows): a
    generator that yields ``(coherence (F, Ne, Nm) float32 numpy, freqs)`` per ``(eeg, emg)`` item, in order.

    Same arithmetic as :func:`welch_magnitude_squared_coherence` item by item, but the items are pipelined over
    three CUDA streams - upload of item i + 1, K1 + K2 of item i and download of item i - 1 overlap - so a sweep runs
    at the speed of the PCIe upload instead of the sum of the three.  Items must be time-first
    ``(n_samples, n_channels)``; pinned float32 host tensors (``torch.Tensor.pin_memory()``) upload asynchronously,
    anything else is staged through a pinned buffer first (one extra host pass).  The yielded coherence array stays
    valid while the NEXT item is fetched and is overwritten when the one after that is requested (three buffer
    sets rotate: one uploading, one computing / downloading, one in the caller's hands); copy it to keep it longer."""
    first, it = _first_item(recordings)
    if first is None:
        return
    dev = _device()
    n, ne, nm = int(first[0].shape[0]), int(first[0].shape[1]), int(first[1].shape[1])
    starts_h, win, dmode, lo, hi, freqs = _welch_plan(n, sampling_freq, nperseg, noverlap, window, detrend, freq_band,
                                                       segment_starts)
    F, L = hi - lo + 1, len(starts_h)
    starts_d = torch.as_tensor(starts_h).to(dev)
    wd = torch.from_numpy(win).to(dev)
    is_hann = _is_periodic_hann(win)
    ne_p, nm_p = ne + (ne & 1), nm + (nm & 1)                    # even channel pitch for K2's TMA

    def compute(slot, i):
        if "S" not in slot:
            slot["S"] = torch.empty((L, 1, F, ne_p + nm_p), dtype=torch.complex64, device=dev)
        sp = slot["S"]
        K.welch_spectra_pair(slot["eeg"], slot["emg"], starts_h, starts_d, wd, is_hann, dmode, lo, hi, sp[..., :ne],
                             sp[..., ne_p:ne_p + nm])
        flat = sp.view(L, F, ne_p + nm_p)
        res = K.csd_msc(flat[:, :, :ne], flat[:, :, ne_p:ne_p + nm])
        return {"coherence": res.coh}

    pipe = _RecordingPipeline(n, ne, nm, {"coherence": ((F, ne, nm), torch.float32)}, compute)
    for out in pipe.run(first, it):
        yield out["coherence"], freqs[lo:hi + 1]


def local_neighbor_coherence(data, neighbor_mapping, sampling_freq: float, nperseg: int = 256) -> float:
    """Average magnitude-squared coherence of every electrode with its neighbours - the quantity
    ``BiosignalPreprocessor.validate_spatial_filtering`` evaluates before / after spatial filtering with one
    ``scipy.signal.coherence`` call per neighbour pair ("~2-5 s per electrode", preprocessing.py:1214-1248;
    SURVEY.md 8f row N4).  Here all pairs come from ONE batched K1 + K2 pass.  ``neighbor_mapping[ch]`` lists the
    neighbour indices of channel ch; mean over frequencies, then neighbours, then electrodes (nan-aware).
    The DC bin is numerically 0/0 after the per-segment detrend in both implementations; it carries 1/F of the
    frequency mean."""
    pc = welch_magnitude_squared_coherence(data, data, sampling_freq, nperseg=nperseg)
    coh = pc.coherence
    if isinstance(coh, torch.Tensor):
        coh = coh.cpu().numpy()
    mean_f = np.nanmean(coh, axis=0)                                  # (C, C)
    per_ch = [np.nanmean([mean_f[ch, nb] for nb in nbs]) if len(nbs) else np.nan
              for ch, nbs in enumerate(neighbor_mapping)]
    return float(np.nanmean(per_ch))


# ----------------------------------------------------------------------------- spectrogram files
def save_spectrograms(spectrograms, time_centers, frequencies, modality: str, save_dir: str | Path,
                      identifier_suffix: str = ""):
    """Three timestamped .npy files with the reference's naming, signal_features.py:1033-1046."""
    spectrograms = spectrograms.cpu().numpy() if isinstance(spectrograms, torch.Tensor) else spectrograms
    save_dir = Path(save_dir)
    print(f"Saving {modality} spectrograms of shape {spectrograms.shape} alongside time-centers and "
          f"frequencies to:\n\t{save_dir}")
    time_center_diffs = np.diff(time_centers)
    window_length_sec = np.nanmin(np.where(time_center_diffs > 0, time_center_diffs, np.nan))
    suffix = f" {identifier_suffix}" if identifier_suffix != "" else ""
    for obj, title in [
        (spectrograms, f"{modality} Spectrograms {spectrograms.shape[2]}ch {window_length_sec:.2f}sec_step{suffix}"),
        (time_centers, f"{modality} Timecenters {len(time_centers)}windows{suffix}"),
        (frequencies, f"{modality} Frequencies {len(frequencies)}freqs{suffix}"),
    ]:
        np.save(save_dir / filemgmt.file_title(title, ".npy"), obj)


def fetch_stored_spectrograms(dir: Path | str, modality: str, file_identifier: str | list[str] | None = None,
                              expected_n_channels: int | None = None):
    """Newest matching (spectrograms, timecenters, frequencies), signal_features.py:1050-1100."""
    ids = ([file_identifier] if isinstance(file_identifier, str)
           else file_identifier if file_identifier is not None else [])
    spectrograms = np.load(filemgmt.most_recent_file(dir, ".npy", [f"{modality}", "Spectrograms"] + ids))
    if expected_n_channels is not None and spectrograms.ndim >= 3:
        actual = spectrograms.shape[2]
        if actual != expected_n_channels:
            raise ValueError(
                f"fetch_stored_spectrograms: expected {expected_n_channels} channels "
                f"on axis 2 but loaded spectrogram has {actual} "
                f"(modality={modality!r}, file_identifier={file_identifier!r}). "
                f"Check that the correct file is being loaded.")
    timecenters = np.load(filemgmt.most_recent_file(dir, ".npy", [f"{modality}", "Timecenters"] + ids))
    frequencies = np.load(filemgmt.most_recent_file(dir, ".npy", [f"{modality}", "Frequencies"] + ids))
    return spectrograms, timecenters, frequencies
