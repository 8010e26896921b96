"""Band / channel aggregation of stored spectrograms - host-side glue between feature extraction and the
statistics workflows (reference ``src/pipeline/signal_features.py:1174-1502``).  Operates
on the (n_times, n_freqs, n_channels) arrays written by ``save_spectrograms``; plain numpy, no device work."""
from __future__ import annotations

from typing import Literal

import numpy as np

# signal_features.py:1446-1455 (note: differs from the module-level FREQUENCY_BANDS)
PSD_FREQUENCY_BANDS = {
    'all': (0, 250), 'slow': (0, 40), 'fast': (60, 250), 'delta': (0.5, 4), 'theta': (4, 8), 'alpha': (8, 12),
    'beta': (13, 30), 'gamma': (30, 100),
}


def aggregate_psd_spectrogram(psd_spectrograms: np.ndarray, psd_freqs: np.ndarray = None, normalize_mvc: bool = False,
                              is_log_scaled: bool = False, freq_slice: tuple[float, float] | str = None,
                              channel_indices: list[int] = None,
                              aggregation_ops: list[tuple[str, int]] = None) -> np.ndarray:
    """MVC normalisation -> frequency slice (``lo <= f <= hi``, both edges INCLUSIVE) -> channel slice ->
    sequential nanmean / nanmax reductions, signal_features.py:1374-1502."""
    out = psd_spectrograms.copy()
    if normalize_mvc and not is_log_scaled:
        peak = out.max(axis=0, keepdims=True).max(axis=1, keepdims=True)     # per channel over time and frequency
        out = out / peak * 100
    if freq_slice is not None:
        if psd_freqs is None:
            raise ValueError("psd_freqs must be provided when using freq_slice")
        if isinstance(freq_slice, str):
            if freq_slice not in PSD_FREQUENCY_BANDS:
                raise ValueError(f"Unknown frequency band '{freq_slice}'. "
                                 f"Available bands: {', '.join(PSD_FREQUENCY_BANDS.keys())}")
            lo, hi = PSD_FREQUENCY_BANDS[freq_slice]
        else:
            lo, hi = freq_slice
        out = out[:, (psd_freqs >= lo) & (psd_freqs <= hi), :]
    if channel_indices is not None:
        out = out[:, :, channel_indices]
    for op, axis in (aggregation_ops or ()):
        if op == 'mean':
            out = np.nanmean(out, axis=axis)
        elif op == 'max':
            out = np.nanmax(out, axis=axis)
        else:
            raise ValueError(f"Unknown operator '{op}'. Supported operators: 'mean', 'max'")
    return out


def aggregate_spectrogram_over_frequency_band(
        spectrograms: np.ndarray, freqs: np.ndarray, behaviour: Literal['max', 'mean'] = 'mean',
        frequency_bands: dict | None = None, log_transform: bool = False, log_epsilon: float = 1e-10,
        frequency_axis: int = 1, pre_aggregate_axis: tuple[int, Literal['max', 'mean']] | None = None,
        lower_array: np.ndarray | None = None, upper_array: np.ndarray | None = None, *,
        strict_band_mask: bool = False, default_bands: dict | None = None):
    """Per-band reduction over the frequency axis with optional coherent CI bounds, signal_features.py:1174-1371.
    Band membership is ``lo <= f < hi`` (upper edge EXCLUSIVE - unlike ``aggregate_psd_spectrogram``).

    Fidelity note: the reference selects the band with ``np.take(spectrograms, frequency_mask, axis=...)``
    (:1303, :1323-1324, :1340-1341).  ``np.take`` does not mask: it converts the boolean mask to the INDICES
    0 / 1, so the "subset" has len(freqs) entries drawn from the first two frequency bins only.  The default
    reproduces that behaviour bit for bit (a drop-in must return what the reference returns);
    ``strict_band_mask=True`` applies the boolean mask that was evidently intended."""
    bands = frequency_bands if frequency_bands is not None else default_bands
    if spectrograms.ndim < 2 + int(pre_aggregate_axis is not None):
        raise ValueError(f"spectrograms must have at least {2 + int(pre_aggregate_axis is not None)} dimensions, "
                         f"got shape {spectrograms.shape}")
    for name, arr in (("lower_array", lower_array), ("upper_array", upper_array)):
        if arr is not None and arr.shape != spectrograms.shape:
            raise ValueError(f"{name} shape {arr.shape} must match spectrograms shape {spectrograms.shape}")
    if (lower_array is None) != (upper_array is None):
        raise ValueError("lower_array and upper_array must both be provided or both be None")
    has_bounds = lower_array is not None
    if len(freqs) != spectrograms.shape[frequency_axis]:
        raise ValueError(f"freqs length ({len(freqs)}) must match spectrograms frequency axis "
                         f"({spectrograms.shape[frequency_axis]})")
    if not bands:
        raise ValueError("frequency_bands dict cannot be empty")
    arrays = [spectrograms] + ([lower_array, upper_array] if has_bounds else [])
    if pre_aggregate_axis is not None:
        ax, how = pre_aggregate_axis
        if how not in ('max', 'mean'):
            raise ValueError(f"Unknown behavior for pre_aggregate_axis '{how}'")
        red = np.max if how == 'max' else np.mean
        arrays = [red(a, axis=ax, keepdims=True) for a in arrays]
    squeeze_axes = (frequency_axis,) if pre_aggregate_axis is None else (frequency_axis, pre_aggregate_axis[0])
    result = {}
    for label, (lo, hi) in bands.items():
        if lo < freqs.min() or hi > freqs.max():
            raise ValueError(f"Band '{label}' range ({lo}, {hi}) exceeds available frequencies "
                             f"({freqs.min():.2f}, {freqs.max():.2f})")
        mask = (freqs >= lo) & (freqs < hi)
        if not mask.any():
            print(f"No frequencies found for band '{label}' in range ({lo}, {hi})")
        sel = np.flatnonzero(mask) if strict_band_mask else mask.astype(np.intp)
        subs = [np.take(a, sel, axis=frequency_axis) for a in arrays]
        if log_transform:
            subs[0] = np.log10(subs[0] + log_epsilon)
        if behaviour == 'max':
            idx = np.argmax(subs[0], axis=frequency_axis, keepdims=True)
            outs = [np.take_along_axis(a, idx, axis=frequency_axis) for a in subs]
        elif behaviour == 'mean':
            outs = [np.mean(a, axis=frequency_axis, keepdims=True) for a in subs]
        else:
            raise ValueError(f"Unknown behaviour '{behaviour}'")
        outs = [np.squeeze(a, axis=squeeze_axes) for a in outs]
        result[label] = tuple(outs) if has_bounds else outs[0]
    return result
