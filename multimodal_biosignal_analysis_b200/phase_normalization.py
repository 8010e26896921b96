"""Force-cycle phase normalisation - the CBPA front-end that turns per-window band power of one trial into
cycle-wise profiles on a common 0-360 degree grid (reference ``src/pipeline/data_analysis.py:960-1233``,
SURVEY.md 8f row N2).  A trial holds a few dozen windows, so this is host-side numpy; the arithmetic follows the
reference step by step (same clipping, duplicate-phase averaging, wrap padding, scipy interp1d) so that results
are identical.  The reference's debug plotting switches are not carried over."""
from __future__ import annotations

from typing import Literal

import numpy as np
from scipy.interpolate import interp1d

_EPS = 1e-9


def _average_duplicates(phase: np.ndarray, values: np.ndarray):
    """Mean of the samples that share a phase value (per channel), phases ascending."""
    uniq, inverse, counts = np.unique(phase, return_inverse=True, return_counts=True)
    if values.ndim == 1:
        return uniq, np.bincount(inverse, weights=values) / counts
    avg = np.zeros((len(uniq), values.shape[1]), dtype=float)
    for ch in range(values.shape[1]):
        avg[:, ch] = np.bincount(inverse, weights=values[:, ch]) / counts
    return uniq, avg


def _nearest_bin_profile(phase: np.ndarray, values: np.ndarray, grid: np.ndarray) -> np.ndarray:
    """Mean of the samples assigned to their circularly nearest grid bin; empty bins are NaN."""
    dist = np.abs(phase[:, None] - grid[None, :])
    nearest = np.minimum(dist, 360.0 - dist).argmin(axis=1)
    counts = np.bincount(nearest, minlength=len(grid))
    filled = counts > 0
    if values.ndim == 1:
        profile = np.full(len(grid), np.nan)
        sums = np.bincount(nearest, weights=values, minlength=len(grid))
        profile[filled] = sums[filled] / counts[filled]
        return profile
    profile = np.full((len(grid), values.shape[1]), np.nan)
    for ch in range(values.shape[1]):
        sums = np.bincount(nearest, weights=values[:, ch], minlength=len(grid))
        profile[filled, ch] = sums[filled] / counts[filled]
    return profile


def phase_normalize_cycles(signal: np.ndarray, t_rel: np.ndarray, task_freq: float, trial_dur_sec: float,
                           phase_grid: np.ndarray, min_samples_per_cycle: int, start_offset_sec: float = 0.0,
                           min_cycle_coverage_ratio: float = 0.8, use_interpolation: bool = True,
                           interpolation_kind: Literal['linear', 'nearest'] = 'linear',
                           show_debug_trial_wise_plots: bool = False, show_debug_cycle_wise_plots: bool = False,
                           phase_wraparound_coverage_threshold: float = 0.8, verbose: bool = True) -> list[np.ndarray]:
    """One profile per accepted task cycle: (len(phase_grid),) for 1-D signals, (len(phase_grid), n_channels) for
    (n_samples, n_channels) signals.  Cycles with fewer than ``min_samples_per_cycle`` samples or a phase coverage
    below ``min_cycle_coverage_ratio`` are skipped; near-complete cycles (coverage >= the wrap-around threshold) are
    padded with copies from the opposite end of the phase axis before interpolation; a closed phase grid
    (last bin = first bin + 360) gets ``profile[-1] = profile[0]``."""
    if show_debug_trial_wise_plots or show_debug_cycle_wise_plots:
        raise NotImplementedError("debug plots need the reference's matplotlib environment")
    if not (0.0 <= float(min_cycle_coverage_ratio) <= 1.0):
        raise ValueError("min_cycle_coverage_ratio must be within [0, 1].")
    if use_interpolation and interpolation_kind not in {'linear', 'nearest'}:
        raise ValueError("interpolation_kind must be 'linear' or 'nearest'.")
    values = np.asarray(signal)
    times = np.asarray(t_rel, dtype=float)
    if values.shape[0] != times.shape[0]:
        raise ValueError("signal and t_rel must have the same length along axis 0.")
    if task_freq <= 0 or values.shape[0] < min_samples_per_cycle:
        return []
    cycle_dur = 1.0 / task_freq
    first_cycle = int(np.floor(start_offset_sec * task_freq))
    n_cycles = int(np.floor(trial_dur_sec * task_freq + _EPS))
    grid = np.asarray(phase_grid, dtype=float)
    if n_cycles <= 0 or grid.size == 0:
        return []
    closed_axis = len(grid) >= 2 and bool(np.isclose(np.mod(grid - grid[0], 360.0)[-1], 0.0, atol=_EPS))
    by_time = np.argsort(times)
    times, values = times[by_time], values[by_time]

    profiles: list[np.ndarray] = []
    for c in range(first_cycle, n_cycles):
        t0, t1 = c * cycle_dur, (c + 1) * cycle_dur
        inside = (times >= t0) & (times < t1)
        if int(inside.sum()) < min_samples_per_cycle:
            continue
        phase = np.clip(((times[inside] - t0) / cycle_dur) * 360.0, 0.0, 360.0 - _EPS)
        vals = values[inside]
        by_phase = np.argsort(phase)
        phase, vals = phase[by_phase], vals[by_phase]
        coverage = (phase[-1] - phase[0]) / 360.0
        if coverage < min_cycle_coverage_ratio:
            continue
        if use_interpolation:
            uniq, avg = _average_duplicates(phase, vals)
            if uniq.size < 2:
                continue
            if verbose and phase_wraparound_coverage_threshold > min_cycle_coverage_ratio:
                print("phase_normalize_cycles [WARNING] min_cycle_coverage_ratio="
                      f"{min_cycle_coverage_ratio:.2f} < phase_wraparound_coverage_threshold="
                      f"{phase_wraparound_coverage_threshold:.2f}: cycles in between are kept without wrap-around "
                      "padding and may have NaN boundary bins.")
            if coverage >= phase_wraparound_coverage_threshold:
                pad = max(1, len(uniq) // 4)
                xp = np.concatenate([uniq[-pad:] - 360.0, uniq, uniq[:pad] + 360.0])
                fp = np.concatenate([avg[-pad:], avg, avg[:pad]], axis=0)
            else:
                xp, fp = uniq, avg
            profile = interp1d(xp, fp, kind=interpolation_kind, axis=0, bounds_error=False, fill_value=np.nan,
                               assume_sorted=True)(grid)
            profile = np.asarray(profile, dtype=float)
            if vals.ndim != 1:
                profile = profile.reshape(len(grid), -1)
        else:
            profile = _nearest_bin_profile(phase, vals, grid)
        if closed_axis:
            profile[-1] = profile[0]
        profiles.append(profile)
    return profiles
