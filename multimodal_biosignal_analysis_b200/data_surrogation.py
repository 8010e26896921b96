"""Drop-in for the reference's ``src/pipeline/data_surrogation.py`` plus the surrogate-null API.

The reference module only holds test-data corrupters (``insert_bad_channels`` :19-65,
``add_noise_to_channels`` :69-148, ``generate_noise`` :151-199); they are tiny host-side numpy
helpers driven by the global ``random`` / ``np.random`` state and are kept with identical
signatures and behaviour.  The statistical surrogates are NEW (the reference's stand-in is the
analytic Beta threshold, signal_features.py:470-481): they reuse the whitened spectra cached by
``signal_features.pooled_coherence`` so that a surrogate costs one cross-spectral contraction, shard
across ranks by global surrogate index and exchange only counts and per-surrogate maxima.
"""
from __future__ import annotations

import random
from typing import Literal

import numpy as np
import torch

from . import dist as cdist
from . import kernels as K


def check_2d_numpy_array(input_array: np.ndarray, axis: Literal[0, 1] = None):
    if len(input_array.shape) == 1:
        input_array = input_array[:, np.newaxis]
        if axis is None:
            axis = 0
    else:
        if axis is None:
            raise AttributeError("For 2D signal arrays, axis needs to be defined!")
    return input_array, axis


def insert_bad_channels(input_array: np.ndarray, axis: Literal[0, 1] = None, n_channels: int = 5,
                        scale_range: tuple[float, float] = (10.0, 15.0)) -> tuple[np.ndarray, list[int]]:
    """Scale ``n_channels`` randomly chosen columns by U(scale_range); returns the copy and the
    1-based indices of the amended channels (data_surrogation.py:19-65).  As in the reference,
    channel 0 is never drawn (``range(1, ...)``) and the count comes from ``shape[axis + 1 % 2]``."""
    input_array, axis = check_2d_numpy_array(input_array, axis)
    output_array = input_array.copy()
    lo, hi = scale_range
    picked = random.sample(range(1, input_array.shape[axis + 1 % 2]), k=n_channels)
    amended = []
    for ch in picked:
        factor = lo + np.random.rand() * (hi - lo)
        output_array[:, ch] = input_array[:, ch] * factor
        amended.append(ch + 1)
    return output_array, amended


def generate_noise(shape: tuple, noise_type: str, amplitude: float) -> np.ndarray:
    """White or pink (white rFFT scaled by 1/sqrt(f)) noise with RMS ``amplitude``
    (data_surrogation.py:151-199)."""
    if noise_type == "white":
        noise = np.random.normal(0, 1, shape)
    elif noise_type == "pink":
        n = shape[0]
        spectrum = np.fft.rfft(np.random.normal(0, 1, n))
        f = np.fft.rfftfreq(n)
        f[0] = 1
        noise = np.fft.irfft(spectrum / np.sqrt(f), n=n)
        if len(shape) > 1:
            noise = np.tile(noise[:, np.newaxis], (1, shape[1]))
    else:
        raise ValueError(f"Unknown noise_type: {noise_type}")
    return noise * (amplitude / np.sqrt(np.mean(noise ** 2)))


def add_noise_to_channels(input_array: np.ndarray, noise_db: float, channels: list[int],
                          axis: Literal[0, 1] = 0, noise_type: Literal["white", "pink"] = "white",
                          random_seed: int = None) -> np.ndarray:
    """Add noise at a target SNR (dB) to the listed channels (data_surrogation.py:69-148)."""
    if random_seed is not None:
        np.random.seed(random_seed)
    array, axis = check_2d_numpy_array(input_array, axis)
    max_channels = array.shape[1 - axis]
    if not all(0 <= ch < max_channels for ch in channels):
        raise ValueError(f"Channel indices must be in range [0, {max_channels - 1}]")
    noisy = array.copy()
    for ch in channels:
        sig = noisy[:, ch] if axis == 0 else noisy[ch, :]
        noise_rms = np.sqrt(np.mean(sig ** 2) / 10 ** (noise_db / 10))
        noise = generate_noise(sig.shape, noise_type, noise_rms)
        if axis == 0:
            noisy[:, ch] = sig + noise
        else:
            noisy[ch, :] = sig + noise
    return noisy


# ----------------------------------------------------------------------------- surrogate null (new)
def _plan(n_surrogates: int, n_freqs: int, shard: str):
    """Work split of one null over the ranks: ``(s_begin, s_end, f_range, by_frequency)``.

    "frequency" (default when every rank can get a bin): each rank runs ALL surrogates on its own slice of the
    frequency axis, so operand generation, phases and contraction all shrink with the rank count (a shift null
    costs at most n_positions - 1 passes however the surrogates are split, so only this split scales it);
    "surrogate": each rank runs a slice of the surrogate index on all bins.  Indices are global either way, the
    null does not depend on the split."""
    rank, world = cdist.world()
    if shard not in ("auto", "frequency", "surrogate"):
        raise ValueError("shard must be 'auto', 'frequency' or 'surrogate'")
    by_freq = world > 1 and (shard == "frequency" or (shard == "auto" and n_freqs >= world))
    if by_freq:
        return 0, n_surrogates, cdist.shard_range(n_freqs, rank, world), True
    begin, end = cdist.shard_range(n_surrogates, rank, world)
    return begin, end, None, False


def _finish(pooled, exceed_d: torch.Tensor, max_local: torch.Tensor, n_surrogates: int, alpha: float,
            by_frequency: bool = False, f_range=None):
    if by_frequency:
        # ranks own disjoint bin slices and all surrogates of them: ONE all-gather moves every rank's slice of the
        # counts together with its per-surrogate maxima (instead of an all-reduce over the whole (F, Ne, Nm) array
        # plus a second max-reduction)
        exceed_d, max_stat = cdist.all_gather_frequency_slices(exceed_d, max_local, f_range)
    else:
        cdist.all_reduce_sum_(exceed_d)                          # disjoint surrogates: counts add up
        max_stat = cdist.all_gather_ranges(max_local, n_surrogates)
    exceed = exceed_d.cpu().numpy().astype(np.int64)
    ms = max_stat.cpu().numpy()
    return {
        "exceed": exceed,                                       # #{s : C_s >= C_obs} per (f, i, j)
        "p_values": (1.0 + exceed) / (1.0 + n_surrogates),
        "max_stat": ms,                                         # max over (f, i, j) per surrogate
        "threshold_fwe": float(np.quantile(ms, 1.0 - alpha)) if n_surrogates else float("nan"),
        "n_surrogates": n_surrogates,
        "coherence": pooled.coherence,
        "freqs": pooled.freqs,
    }


def circular_shift_surrogate_null(pooled, n_surrogates: int = 1000, seed: int | None = 0,
                                  shifts: np.ndarray | None = None, alpha: float = 0.05,
                                  shard: str = "auto") -> dict:
    """Circular time-shift surrogates: surrogate s rotates the EMG segment (window) index by
    ``shifts[s]`` in [1, n_positions - 1] (whole windows for multitaper pooling, tapers stay
    aligned).  ``shifts`` is a host-supplied int32 table; by default it is drawn from
    ``np.random.default_rng(seed)``.  Returns exceedance counts, p-values, the per-surrogate
    max statistic and its (1 - alpha) quantile (family-wise threshold)."""
    csd = pooled.device_result
    L = csd.dims[0]
    n_pos = L // pooled.group
    if n_pos < 2:
        raise ValueError("circular shift surrogates need at least two segments")
    if shifts is None:
        shifts = np.random.default_rng(seed).integers(1, n_pos, n_surrogates)
    shifts = np.ascontiguousarray(shifts, dtype=np.int32)
    if shifts.shape != (n_surrogates,):
        raise ValueError("one shift per surrogate")
    # only n_pos - 1 distinct surrogates exist: the smallest p-value the null can resolve is 1 / n_pos however many
    # are drawn (the device contracts every distinct shift once and weights it with its multiplicity)
    n_distinct = int(np.unique(shifts % n_pos).size)
    if n_surrogates > n_pos - 1:
        import warnings
        warnings.warn(f"circular-shift null: {n_surrogates} surrogates drawn from only {n_pos - 1} distinct shifts "
                      f"({n_distinct} used): p-values below {1.0 / n_pos:.3g} cannot be resolved; use the "
                      "phase-randomised null for finer p-values", RuntimeWarning, stacklevel=2)
    begin, end, f_range, by_freq = _plan(n_surrogates, csd.dims[1], shard)
    dev = csd.coh.device
    exceed, max_local = K.surrogate_null(csd, K.SURR_SHIFT, begin, end,
                                         shifts=torch.from_numpy(shifts[begin:end]).to(dev), group=pooled.group,
                                         f_range=f_range)
    out = _finish(pooled, exceed, max_local, n_surrogates, alpha, by_freq, f_range)
    out["n_distinct_shifts"] = n_distinct
    out["p_resolution"] = 1.0 / n_pos
    return out


def phase_randomised_surrogate_null(pooled, n_surrogates: int = 1000, seed: int = 0, alpha: float = 0.05,
                                    shard: str = "auto", thresholds: bool = False, threshold_passes: int = 3,
                                    hist_bins: int = 128, return_hist: bool = False,
                                    hist_range: tuple[float, float] = (0.0, 1.0)) -> dict:
    """Phase-randomised surrogates: every EMG spectrum is rotated by one random phase per
    (surrogate, segment, frequency), shared by all EMG channels; phases come from
    Philox4x32-10(seed; s, l, f) so any sharding of the surrogate index gives the same null.

    ``thresholds=True`` adds the per-pair significance thresholds of BASELINE config 3: ``threshold[f, i, j]`` =
    the (1 - alpha) quantile (order statistic, numpy ``method="higher"``) of the pair's own null coherences, from
    per-pair null histograms accumulated on the device (:func:`null_quantile_thresholds`), and ``significant = coherence >
    threshold`` - the surrogate counterpart of the reference's analytic ``apply_threshold_filtering``
    (signal_features.py:581-604).  ``return_hist=True`` also returns every pair's null histogram over ``hist_range``
    (``null_hist`` (F, Ne, Nm, hist_bins) uniform coherence bins, ``null_hist_edges``, ``null_hist_below`` = the
    surrogates under the range; those over it are not counted)."""
    csd = pooled.device_result
    begin, end, f_range, by_freq = _plan(n_surrogates, csd.dims[1], shard)
    exceed, max_local = K.surrogate_null(csd, K.SURR_PHASE, begin, end, seed=seed, f_range=f_range,
                                         keep_phase_operands=(thresholds or return_hist) and (begin, end) == (0, n_surrogates))
    out = _finish(pooled, exceed, max_local, n_surrogates, alpha, by_freq, f_range)
    if thresholds:
        thr, _ = null_quantile_thresholds(csd, n_surrogates, seed, 1.0 - alpha, passes=threshold_passes, shard=shard)
        thr_h = thr.cpu().numpy().astype(np.float64)
        out["threshold"] = thr_h
        coh = pooled.coherence
        out["significant"] = (coh.cpu().numpy() if isinstance(coh, torch.Tensor) else coh) > thr_h
    if return_hist:
        lo_h, hi_h = float(hist_range[0]), float(hist_range[1])
        F, Ne, Nm = csd.dims[1:]
        fb, fe = f_range if by_freq else (0, F)
        lo_t = torch.full((F, Ne, Nm), lo_h, dtype=torch.float32, device=csd.coh.device)
        sc_t = torch.full((F, Ne, Nm), hist_bins / (hi_h - lo_h), dtype=torch.float32, device=csd.coh.device)
        hist, below = K.surrogate_null_hist(csd, 0, n_surrogates, seed=seed, n_bins=hist_bins, bin_lo=lo_t,
                                            bin_scale=sc_t, f_range=(fb, fe))
        if by_freq:
            cdist.all_reduce_sum_(hist)
            cdist.all_reduce_sum_(below)
        out["null_hist"] = hist.cpu().numpy()
        out["null_hist_below"] = below.cpu().numpy()
        out["null_hist_edges"] = np.linspace(lo_h, hi_h, hist_bins + 1)
    return out


def null_quantile_thresholds(csd, n_surrogates: int, seed: int, q: float, passes: int = 3, n_bins: int = 128,
                             shard: str = "auto"):
    """Per-pair q-quantile of the phase-surrogate null for every (f, i, j) without ever materialising the
    (n_surrogates, F, Ne, Nm) stack.  The quantile is an ORDER STATISTIC of the pair's surrogate coherences,
    ``np.quantile(C_s[:, f, i, j], q, method="higher")`` = the k-th smallest with k = ceil(q (n - 1)) - the
    conservative choice for a significance threshold (at most (1 - q) n surrogates lie above it) and an actually
    observed null value.

    Every pass re-runs the null GEMM with a per-pair window of ``n_bins`` bins (``cmc_surrogate_null_hist``):
    surrogates under the window are only counted, those inside are histogrammed, ``cmc_hist_select`` finds the bin
    of rank k and the next pass zooms into that bin (or into the part of [0, 1] under / over the window when the
    rank fell outside).  The first window is placed from the analytic mean m of the null, E[C_s] = sum_l |Z_l|^2
    (an approximately exponential null has its q-quantile at -m ln(1 - q)): [0.5, 0.5 + 16 / -ln(1 - q)] x that
    value, so that ~80 % of the surrogates stay under it.  After p passes a well-placed pair is resolved to
    16 m / n_bins^p (3 passes, 128 bins, m = 0.005: 4e-8), a misplaced one to n_bins^-(p - 1).  Multi-rank: ranks
    split the frequency axis and the threshold slices are summed.  Returns (threshold float32 (F, Ne, Nm) CUDA
    tensor, dict with the last window: lo, width, hist, below)."""
    if passes < 1:
        raise ValueError("passes must be >= 1")
    L, F, Ne, Nm = csd.dims
    dev = csd.coh.device
    n = int(n_surrogates)
    if n < 1:
        raise ValueError("n_surrogates must be >= 1")
    if not 0.0 < q < 1.0:
        raise ValueError("q must lie in (0, 1)")
    _, _, f_range, by_freq = _plan(n, F, shard if shard != "surrogate" else "frequency")
    fb, fe = f_range if by_freq else (0, F)
    k = min(int(np.ceil(q * (n - 1) - 1e-9)), n - 1)        # 0-based rank of the order statistic
    m = csd.null_mean()
    qhat = m * float(-np.log1p(-q))
    lo = (0.5 * qhat).clamp(max=1.0)
    width = (16.0 * m).clamp(min=1e-12)
    width = torch.minimum(width, (1.0 - lo).clamp(min=1e-12))
    for p in range(passes):
        scale = (float(n_bins) / width).contiguous()
        hist, below = K.surrogate_null_hist(csd, 0, n, seed=seed, n_bins=n_bins, bin_lo=lo.contiguous(),
                                            bin_scale=scale, f_range=(fb, fe), keep_operands=p + 1 < passes)
        last = dict(lo=lo, width=width, hist=hist, below=below.clone())
        b = K.hist_select(hist, k, below)                    # bin of rank k; below <- count under that bin
        bw = width / n_bins
        under, over = b < 0, b >= n_bins
        x = torch.where(under, lo, torch.where(over, lo + width, lo + (b.to(torch.float32) + 0.5) * bw))
        if p + 1 < passes:
            new_lo = torch.where(under, torch.zeros_like(lo), torch.where(over, lo + width, lo + b.to(torch.float32) * bw))
            new_w = torch.where(under, lo, torch.where(over, 1.0 - (lo + width), bw))
            lo, width = new_lo.clamp(0.0, 1.0), new_w.clamp(min=1e-12)
    thr = x.clamp(0.0, 1.0)
    if by_freq:
        mask = torch.zeros(F, dtype=torch.bool, device=dev)
        mask[fb:fe] = True
        thr = torch.where(mask[:, None, None], thr, torch.zeros_like(thr))
        cdist.all_reduce_sum_(thr)
    return thr, last


def surrogate_null_sweep(recordings, sampling_freq: float, nperseg: int = 256, noverlap: int | None = None,
                         window: str = "hann", detrend: str | bool = "constant",
                         freq_band: tuple[float, float] | None = None, segment_starts=None,
                         n_surrogates: int = 1000, mode: str = "phase", seed: int = 0, alpha: float = 0.05,
                         unit_indices=None):
    """Welch coherence + surrogate null of MANY recordings of equal shape - the subject-condition sweep of BASELINE
    config 5 (reference loop: ``src/subject_feature_extraction_workflow.py:37``, per-subject CMC call ``:246-255``).

    Generator: per ``(eeg, emg)`` item, in order, a dict with ``coherence`` (F, Ne, Nm) float32, ``exceed``
    (#{s : C_s >= C_obs}, int32), ``p_values`` (float64), ``max_stat`` (n_surrogates,) float32, ``threshold_fwe``
    (the (1 - alpha) quantile of ``max_stat``), ``freqs`` and ``unit`` (the item's global index).  Per item the device
    runs K1 (both modalities) -> K2 -> operand planes -> the whole null; upload of item i + 1 and download of item
    i - 1 overlap with it on separate streams (``signal_features._RecordingPipeline``), so a sweep is GPU-bound as
    long as one null outlasts one upload.  Arrays stay valid while the next item is fetched (copy to keep longer).

    Item u draws its surrogates from ``seed + unit_indices[u]`` (default: its position in ``recordings``), so a
    rank that is handed ``recordings[rank::world]`` with ``unit_indices=range(rank, n, world)`` reproduces exactly
    what a single process computes for those items - units shard over ranks with NO collective."""
    from . import signal_features as sf
    if mode not in ("phase", "shift"):
        raise ValueError("mode must be 'phase' or 'shift'")
    first, it = sf._first_item(recordings)
    if first is None:
        return
    dev = sf._device()
    n, ne, nm = int(first[0].shape[0]), int(first[0].shape[1]), int(first[1].shape[1])
    starts_h, win, dmode, lo, hi, freqs = sf._welch_plan(n, sampling_freq, nperseg, noverlap, window, detrend,
                                                         freq_band, segment_starts)
    F, L = hi - lo + 1, len(starts_h)
    if mode == "shift" and L < 2:
        raise ValueError("circular shift surrogates need at least two segments")
    starts_d = torch.as_tensor(starts_h).to(dev)
    wd = torch.from_numpy(win).to(dev)
    is_hann = sf._is_periodic_hann(win)
    ne_p, nm_p = ne + (ne & 1), nm + (nm & 1)
    units = iter(unit_indices) if unit_indices is not None else None
    unit_of = {}

    def compute(slot, i):
        u = int(next(units)) if units is not None else i
        unit_of[i] = u
        if "S" not in slot:
            slot["S"] = torch.empty((L, 1, F, ne_p + nm_p), dtype=torch.complex64, device=dev)
        sp = slot["S"]
        K.welch_spectra_pair(slot["eeg"], slot["emg"], starts_h, starts_d, wd, is_hann, dmode, lo, hi, sp[..., :ne],
                             sp[..., ne_p:ne_p + nm])
        flat = sp.view(L, F, ne_p + nm_p)
        csd = K.csd_msc(flat[:, :, :ne], flat[:, :, ne_p:ne_p + nm])
        if mode == "phase":
            exceed, max_stat = K.surrogate_null(csd, K.SURR_PHASE, 0, n_surrogates, seed=seed + u)
        else:
            sh = np.random.default_rng(seed + u).integers(1, L, n_surrogates).astype(np.int32)
            exceed, max_stat = K.surrogate_null(csd, K.SURR_SHIFT, 0, n_surrogates,
                                                shifts=torch.from_numpy(sh).to(dev, non_blocking=True))
        return {"coherence": csd.coh, "exceed": exceed, "max_stat": max_stat}

    pipe = sf._RecordingPipeline(n, ne, nm, {"coherence": ((F, ne, nm), torch.float32),
                                             "exceed": ((F, ne, nm), torch.int32),
                                             "max_stat": ((n_surrogates,), torch.float32)}, compute)
    for k, out in enumerate(pipe.run(first, it)):
        ms = out["max_stat"]
        yield {"coherence": out["coherence"], "exceed": out["exceed"],
               "p_values": (1.0 + out["exceed"]) / (1.0 + n_surrogates), "max_stat": ms,
               "threshold_fwe": float(np.quantile(ms, 1.0 - alpha)) if n_surrogates else float("nan"),
               "n_surrogates": n_surrogates, "freqs": freqs[lo:hi + 1], "unit": unit_of[k]}
