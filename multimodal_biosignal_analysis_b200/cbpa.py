"""Drop-in for the cluster-based permutation analysis of the reference's ``src/pipeline/cbpa.py``.

The reference delegates the arithmetic to ``mne.stats.spatio_temporal_cluster_1samp_test`` /
``permutation_cluster_1samp_test`` (cbpa.py:1027-1042).  This module keeps ``CBPAConfig``,
``run_cbpa`` / ``run_batch``, the adjacency builders and the result-dict contract
(cbpa.py:1051-1056), and provides the two MNE-signature functions on top of the sm_100a CBPA
kernels (``cmc_cbpa_observed`` / ``cmc_cbpa_permute``).  No MNE, no CPU fallback.

Determinism: sign flips come from a host-side table (``make_sign_table``; pass ``signs=`` to supply
your own), never from a library RNG stream; permutations shard across ranks by global index
(``dist.shard_range``) and only the H0 slices are gathered.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from pathlib import Path
from typing import Literal, Optional

import numpy as np
import torch
from scipy import sparse
from scipy.stats import t as t_dist

from . import dist as cdist
from . import file_management as filemgmt
from . import kernels as K
from .channel_layout import EEG_CHANNEL_IND_DICT

EEG_CHANNELS: list[str] = list(EEG_CHANNEL_IND_DICT.keys())
EEG_SFREQ: float = 2048

# cbpa.py:36-43 - must match the feature-extraction workflow
CMC_EEG_CHANNEL_SUBSET: list[str] = ["C5", "C3", "C1", "FC5", "FC3", "FC1", "F3", "CP5", "CP3", "CP1", "P3"]
CMC_CHANNEL_FILE_SUFFIX: str = f"Channels_{'_'.join(CMC_EEG_CHANNEL_SUBSET)}"


@dataclass
class CBPAConfig:
    """One CBPA run - same fields and defaults as the reference dataclass (cbpa.py:50-193)."""
    # Feature
    modality: Literal["PSD", "CMC"] = "PSD"
    modality_file_id: str = "eeg"
    freq_band: str = "alpha"
    channels: Optional[list[str]] = None
    # Contrast
    condition_column: str = "Category or Silence"
    condition_A: str = "Happy"
    condition_B: str = "Silence"
    # Segmentation
    n_within_trial_segs: int = 1
    # Subject subset
    exclude_subjects: list[int] = None
    # CBPA
    alpha_cluster_forming: float = 0.05
    n_permutations: int = 1000
    tail: Literal[-1, 0, 1] = 0
    use_spatio_temporal: bool = True
    n_jobs: int = -1
    seed: int = 42
    # I/O
    data_root: Path = field(default_factory=lambda: Path().resolve().parent)
    psd_time_window_sec: float = 0.25
    cmc_time_window_sec: float = 2.0
    overlap_ratio: float = .5
    psd_is_log_scaled: bool = True
    output_dir: Path = field(
        default_factory=lambda: Path().resolve().parent / "output" / "statistics_post_hoc_testing")
    hypothesis_label: str = "cbpa_run"
    save_plots: bool = True
    show_plots: bool = False
    # Phase normalisation (CMC only)
    use_phase_normalization: bool = False
    n_phase_bins: int = 36
    min_samples_per_cycle: int = 2
    min_cycles_per_condition: int = 3
    # Plot options (consumed by the reference's plotting code only)
    show_target_sine: bool | None = None
    target_sine_min_pct_mvc: float = 7.5
    target_sine_max_pct_mvc: float = 22.5
    target_sine_frequency_hz: float = 0.1
    include_dynamometer_force: bool = True
    phase_start_offset_sec: float | None = None
    force_phase_start_offset_sec: float | None = None
    include_suptitle: bool = False
    use_stretched_window_timestamps: bool = False


# ----------------------------------------------------------------------------- adjacency (host, scipy.sparse)
def combine_adjacency(n_times: int, spatial_adj) -> sparse.csr_matrix:
    """Lattice-in-time x spatial adjacency with index ``t * n_ch + ch`` (what
    ``mne.stats.combine_adjacency(n_times, spatial_adj)`` returns at cbpa.py:237): nodes are
    neighbours when they differ in exactly one of (time by one step, channel by a spatial edge);
    the diagonal is set."""
    sp = sparse.coo_matrix(spatial_adj)
    n_ch = sp.shape[0]
    off = sp.row != sp.col
    srow, scol = sp.row[off].astype(np.int64), sp.col[off].astype(np.int64)
    t = np.arange(n_times, dtype=np.int64)
    ch = np.arange(n_ch, dtype=np.int64)
    lower = (t[:-1, None] * n_ch + ch[None, :]).ravel()
    n = n_times * n_ch
    rows = np.concatenate([(t[:, None] * n_ch + srow[None, :]).ravel(), lower, lower + n_ch, np.arange(n)])
    cols = np.concatenate([(t[:, None] * n_ch + scol[None, :]).ravel(), lower + n_ch, lower, np.arange(n)])
    m = sparse.coo_matrix((np.ones(len(rows)), (rows, cols)), shape=(n, n)).tocsr()
    m.data[:] = 1.0
    return m


def find_ch_adjacency_from_positions(pos_2d: np.ndarray) -> sparse.csr_matrix:
    """Delaunay-triangulation neighbours of 2-D sensor positions - the construction
    ``mne.channels.find_ch_adjacency`` applies to a montage without template (cbpa.py:235)."""
    from scipy.spatial import Delaunay
    tri = Delaunay(np.asarray(pos_2d, dtype=np.float64))
    n = len(pos_2d)
    e = np.concatenate([tri.simplices[:, [a, b]] for a in range(3) for b in range(3) if a != b])
    m = sparse.coo_matrix((np.ones(len(e)), (e[:, 0], e[:, 1])), shape=(n, n)).tocsr()
    m.data[:] = 1.0
    return m


def _build_adjacency(info, n_times: int):
    """cbpa.py:224-243.  ``info`` may be a spatial adjacency matrix, an (n_ch, 2) array of sensor
    positions, or an ``mne.Info`` (only when MNE is installed)."""
    if sparse.issparse(info) or (isinstance(info, np.ndarray) and info.ndim == 2 and info.shape[0] == info.shape[1]):
        spatial_adj = sparse.csr_matrix(info)
    elif isinstance(info, np.ndarray) and info.ndim == 2 and info.shape[1] == 2:
        spatial_adj = find_ch_adjacency_from_positions(info)
    else:
        import mne  # noqa: F401 - optional; only needed for mne.Info inputs
        spatial_adj, _ = mne.channels.find_ch_adjacency(info, ch_type="eeg")
    combined = combine_adjacency(n_times, spatial_adj)
    print(f"  [adjacency] spatial: {spatial_adj.shape}, combined (time×space): {combined.shape}, "
          f"nnz edges: {combined.nnz}")
    return combined


def _add_phase_wraparound(adjacency, n_times: int, n_ch: int, time_grid: np.ndarray):
    """(0, ch) <-> (n_times - 1, ch) edges for circular phase axes, cbpa.py:949-982."""
    ch = np.arange(n_ch)
    first, last = ch, (n_times - 1) * n_ch + ch
    wrap = sparse.coo_matrix((np.ones(2 * n_ch, dtype=bool), (np.r_[first, last], np.r_[last, first])),
                             shape=adjacency.shape).tocsr()
    result = (adjacency.astype(bool) + wrap).astype(bool).tocsr()
    print(f"  [adjacency] Phase wrap-around edges added (0°↔{int(time_grid[-1])}° for {n_ch} channels)")
    return result


# ----------------------------------------------------------------------------- permutation machinery
def make_sign_table(n_permutations: int, n_subjects: int, seed=None, tail: int = 0) -> np.ndarray:
    """int8 (n_rows, n_subjects) table of +-1.  Like MNE the observed ordering counts as one
    permutation, so ``n_permutations - 1`` random rows are drawn; when all
    ``2**(n_subjects - (tail == 0)) - 1`` distinct patterns fit, the test is exact and every
    pattern is enumerated instead.  Two-tailed tests keep subject 0 fixed (t -> -t symmetry)."""
    free = n_subjects - (1 if tail == 0 else 0)
    max_perms = 2 ** free - 1 if free < 62 else np.inf
    if n_permutations - 1 >= max_perms:
        codes = np.arange(1, int(max_perms) + 1, dtype=np.int64)
        bits = (codes[:, None] >> np.arange(free, dtype=np.int64)[None, :]) & 1
    else:
        rng = seed if isinstance(seed, np.random.Generator) else np.random.default_rng(seed)
        bits = rng.integers(0, 2, size=(max(n_permutations - 1, 0), free), dtype=np.int8)
    signs = np.ones((bits.shape[0], n_subjects), dtype=np.int8)
    signs[:, n_subjects - free:] = 1 - 2 * bits.astype(np.int8)
    return signs


def _pvalues(stats_fixed: np.ndarray, h0_fixed: np.ndarray, tail: int) -> np.ndarray:
    """MNE's ``_pval_from_histogram`` on the integer images: exact counts / len(H0), through one sort of H0 and a
    binary search per cluster instead of one pass over H0 per cluster (same values, float64 count / n)."""
    stats_fixed = np.asarray(stats_fixed, dtype=np.int64)
    n = len(h0_fixed)
    if len(stats_fixed) == 0 or n == 0:
        return np.zeros(len(stats_fixed), dtype=np.float64)
    if tail == -1:                                               # #{H0 <= s}
        cnt = np.searchsorted(np.sort(h0_fixed), stats_fixed, side="right")
    elif tail == 1:                                              # #{H0 >= s}
        cnt = n - np.searchsorted(np.sort(h0_fixed), stats_fixed, side="left")
    else:                                                        # #{|H0| >= |s|}
        cnt = n - np.searchsorted(np.sort(np.abs(h0_fixed)), np.abs(stats_fixed), side="left")
    return cnt.astype(np.float64) / float(n)


class DeviceAdjacency:
    """A sparse adjacency already converted to sorted CSR int32 on the current CUDA device.  Pass it as
    ``adjacency=`` to skip the per-call conversion / upload (``sweep.py`` prepares it while the unit stage runs)."""

    def __init__(self, adjacency):
        adj = sparse.csr_matrix(adjacency)
        adj.sort_indices()
        dev = torch.device("cuda", torch.cuda.current_device())
        self.shape = adj.shape
        self.nnz = adj.nnz
        self.indptr = torch.from_numpy(adj.indptr.astype(np.int32)).to(dev)
        self.indices = torch.from_numpy(adj.indices.astype(np.int32)).to(dev)


def permutation_cluster_1samp_test(X, threshold=None, n_permutations: int = 1024, tail: int = 0,
                                   adjacency=None, n_jobs=None, seed=None, out_type: str = "indices",
                                   verbose=None, *, signs: np.ndarray | None = None, return_details: bool = False):
    """GPU equivalent of ``mne.stats.permutation_cluster_1samp_test`` for the options the reference
    uses (sparse adjacency, t_power=1, no TFCE / step-down).  X (n_subj, ...) float64; returns
    ``(t_obs, clusters, cluster_pv, H0)``: t_obs shaped like one observation, clusters as boolean
    masks (``out_type='mask'``) or index tuples, H0 = [observed] + one value per sign-table row.
    ``n_jobs`` / ``verbose`` are accepted for signature compatibility."""
    if adjacency is None:
        raise ValueError("a sparse adjacency is required (the reference always passes one)")
    Xh = X.detach().cpu().numpy() if isinstance(X, torch.Tensor) else np.asarray(X)
    n_subj = Xh.shape[0]
    sample_shape = Xh.shape[1:]
    Xf = np.ascontiguousarray(Xh.reshape(n_subj, -1), dtype=np.float64)
    n_tests = Xf.shape[1]
    if tail not in (-1, 0, 1):
        raise ValueError("tail must be -1, 0 or 1")
    if threshold is None:
        p = 0.05 / (1 + (tail == 0))
        threshold = float(t_dist.ppf(1 - p, n_subj - 1)) * (-1 if tail == -1 else 1)
    if (tail < 0 and threshold > 0) or (tail > 0 and threshold < 0) or (tail == 0 and threshold < 0):
        raise ValueError(f"incompatible tail and threshold signs, got {tail} and {threshold}")
    adj = adjacency if isinstance(adjacency, DeviceAdjacency) else DeviceAdjacency(adjacency)
    if tuple(adj.shape) != (n_tests, n_tests):
        raise ValueError(f"adjacency must be ({n_tests}, {n_tests}), got {tuple(adj.shape)}")
    if signs is None:
        signs = make_sign_table(n_permutations, n_subj, seed, tail)
    signs = np.ascontiguousarray(signs, dtype=np.int8)
    if signs.ndim != 2 or signs.shape[1] != n_subj or not np.all(np.abs(signs) == 1):
        raise ValueError("signs must be an (n_rows, n_subjects) table of +-1")

    dev = torch.device("cuda", torch.cuda.current_device())
    Xd = torch.from_numpy(Xf).to(dev)
    indptr, indices = adj.indptr, adj.indices
    ws = K.cbpa_workspace(Xd)                                    # both calls share the re-tiled copy of X
    t_obs_d, labels_d, mass_d, n_clusters = K.cbpa_observed(Xd, threshold, tail, indptr, indices, ws=ws)
    n_rows = signs.shape[0]
    begin, end = cdist.shard_range(n_rows)
    h0_local = K.cbpa_permute(Xd, torch.from_numpy(signs).to(dev), begin, end, threshold, tail, indptr, indices,
                              ws=ws, tiled=True)
    h0_perm = cdist.all_gather_ranges(h0_local, n_rows).cpu().numpy()
    mass = mass_d.cpu().numpy()
    if n_clusters:
        orig = np.abs(mass).max() if tail == 0 else (mass.max() if tail == 1 else mass.min())
    else:
        orig = 0
    h0_fixed = np.concatenate([[orig], h0_perm]).astype(np.int64)
    cluster_pv = _pvalues(mass, h0_fixed, tail)
    labels = labels_d.cpu().numpy()
    t_obs = t_obs_d.cpu().numpy().reshape(sample_shape)
    clusters = []
    for k in range(1, n_clusters + 1):
        m = labels == k
        clusters.append(m.reshape(sample_shape) if out_type == "mask"
                        else np.unravel_index(np.flatnonzero(m), sample_shape))
    H0 = h0_fixed.astype(np.float64) / K.FIX_SCALE
    if return_details:
        return t_obs, clusters, cluster_pv, H0, dict(labels=labels.reshape(sample_shape), mass_fixed=mass,
                                                     H0_fixed=h0_fixed, signs=signs)
    return t_obs, clusters, cluster_pv, H0


def spatio_temporal_cluster_1samp_test(X, threshold=None, n_permutations: int = 1024, tail: int = 0,
                                       adjacency=None, n_jobs=None, seed=None, out_type: str = "indices",
                                       verbose=None, **kwargs):
    """X (n_subj, n_times, n_ch) with a combined (n_times * n_ch)^2 adjacency - the call the
    reference makes at cbpa.py:1028-1032."""
    if np.ndim(X) != 3:
        raise ValueError("X must be (n_observations, n_times, n_vertices)")
    return permutation_cluster_1samp_test(X, threshold=threshold, n_permutations=n_permutations, tail=tail,
                                          adjacency=adjacency, n_jobs=n_jobs, seed=seed, out_type=out_type,
                                          verbose=verbose, **kwargs)


# ----------------------------------------------------------------------------- contrast front-end (row N2)
def _get_task_freq_for_trial(log_df, t_start, t_end) -> float | None:
    """Modal non-NaN 'Task Frequency' (Hz) inside [t_start, t_end), cbpa.py:250-279."""
    col = log_df.loc[(log_df.index >= t_start) & (log_df.index < t_end), "Task Frequency"].dropna()
    return None if col.empty else float(col.mode().iloc[0])


def _extract_band_power(cfg: CBPAConfig, spectrogram: np.ndarray, freqs: np.ndarray,
                        channel_indices: list[int] | None, freq_pooling: Literal["max", "mean"] = "max",
                        channel_pooling: Literal["max", "mean"] = "max") -> np.ndarray:
    """Stored spectrogram -> (n_windows, n_channels) band power, cbpa.py:564-649: CMC pools the EMG axis (if the
    4-D tensor was stored) and then the band with ``freq_pooling``; PSD takes the band mean."""
    from .spectrogram_aggregation import aggregate_psd_spectrogram
    spec = spectrogram
    if cfg.modality == "CMC":
        if spec.ndim == 4:
            spec = np.nanmean(spec, axis=3) if channel_pooling == "mean" else np.nanmax(spec, axis=3)
        elif spec.ndim != 3:
            raise ValueError(f"Unexpected CMC spectrogram shape {spec.shape}. "
                             "Expected 3D (time,freq,eeg) or 4D (time,freq,eeg,emg).")
    elif spec.ndim != 3:
        raise ValueError(f"Unexpected PSD spectrogram shape {spec.shape}. Expected 3D (time,freq,channel).")
    band_op = freq_pooling if cfg.modality == "CMC" else "mean"
    return aggregate_psd_spectrogram(
        spec, freqs, normalize_mvc=False, channel_indices=channel_indices,
        is_log_scaled=cfg.psd_is_log_scaled if cfg.modality == "PSD" else False, freq_slice=cfg.freq_band,
        aggregation_ops=[(band_op, 1)])


def _band_power_per_phase(cfg: CBPAConfig, band_power: np.ndarray, timestamps, trial_spans: dict, trial_cond_map: dict,
                          log_df, min_cycle_coverage_ratio: float = 0.8) -> dict[str, list[np.ndarray]]:
    """Cycle-wise phase profiles of every trial grouped by condition, cbpa.py:651-725.  ``log_df`` is the enriched
    log frame (DatetimeIndex, 'Task Frequency' column) or a mapping trial_id -> task frequency in Hz."""
    import warnings
    from .phase_normalization import phase_normalize_cycles
    phase_grid = np.linspace(0, 360, cfg.n_phase_bins, endpoint=False)
    out: dict[str, list[np.ndarray]] = {}
    for trial_id, (t_start, t_end) in trial_spans.items():
        condition = trial_cond_map.get(int(trial_id))
        if condition is None:
            continue
        task_freq = (log_df.get(int(trial_id)) if isinstance(log_df, dict)
                     else _get_task_freq_for_trial(log_df, t_start, t_end))
        if task_freq is None or task_freq <= 0:
            warnings.warn(f"  [phase] Trial {trial_id}: Task Frequency missing or zero. Skipping.")
            continue
        step = (cfg.cmc_time_window_sec if cfg.modality == "CMC" else cfg.psd_time_window_sec) * (1.0 - cfg.overlap_ratio)
        samples_per_cycle = (1.0 / task_freq) / step
        if samples_per_cycle < cfg.min_samples_per_cycle:
            warnings.warn(f"  [phase] Trial {trial_id}: only {samples_per_cycle:.1f} CMC samples/cycle at "
                          f"{task_freq} Hz — skipping (min={cfg.min_samples_per_cycle}).")
            continue
        mask = np.asarray((timestamps >= t_start) & (timestamps < t_end))
        if not mask.any():
            continue
        t_rel = np.array([(ts - t_start).total_seconds() for ts in timestamps[mask]], dtype=float)
        offset = float(cfg.phase_start_offset_sec) if cfg.phase_start_offset_sec is not None else float(1.0 / task_freq)
        cycles = phase_normalize_cycles(
            signal=band_power[mask], t_rel=t_rel, task_freq=task_freq,
            trial_dur_sec=(t_end - t_start).total_seconds(), phase_grid=phase_grid,
            min_samples_per_cycle=cfg.min_samples_per_cycle, min_cycle_coverage_ratio=min_cycle_coverage_ratio,
            start_offset_sec=offset)
        out.setdefault(condition, []).extend(cycles)
    return out


def phase_contrast_from_cycles(cfg: CBPAConfig, cycles_by_condition: dict) -> np.ndarray | None:
    """Per-subject A - B profile (n_phase_bins, n_ch): nanmean over the valid cycles of each condition; ``None`` when
    a condition has fewer than ``cfg.min_cycles_per_condition`` cycles (subject skipped), cbpa.py:858-879."""
    a = cycles_by_condition.get(cfg.condition_A, [])
    b = cycles_by_condition.get(cfg.condition_B, [])
    if len(a) < cfg.min_cycles_per_condition or len(b) < cfg.min_cycles_per_condition:
        return None
    return np.nanmean(np.stack(a, axis=0), axis=0) - np.nanmean(np.stack(b, axis=0), axis=0)


# ----------------------------------------------------------------------------- data loading (cbpa.py:282-432)
def _load_subject_data(cfg: CBPAConfig, subject_ind: int):
    """Stored spectrogram + enriched log of one subject (cbpa.py:282-350): ``(spectrogram (n_windows, n_freqs,
    n_channels), freqs, timestamps (tz-aware DatetimeIndex, one per window centre), log_df)``.  Left-handed subjects
    carry the mirrored CMC channel subset in their file names."""
    import pandas as pd
    from . import experiment_log as xlog
    from .signal_features import fetch_stored_spectrograms, mirror_eeg_channel_list
    data = Path(cfg.data_root) / "data"
    feat_dir = data / "precomputed_features" / f"subject_{subject_ind:02}"
    exp_dir = data / "experiment_results" / f"subject_{subject_ind:02}"
    handedness = xlog.fetch_personal_data(exp_dir, False)["Dominant hand"]
    log_df = xlog.fetch_enriched_log_frame(exp_dir, verbose=False)
    log_df.index = xlog.make_timezone_aware(log_df.index)
    qtc_start, qtc_end = xlog.get_qtc_measurement_start_end(log_df, False)
    if cfg.modality == "CMC":
        subset = (mirror_eeg_channel_list(CMC_EEG_CHANNEL_SUBSET, input_is_left=True) if handedness == "Left"
                  else CMC_EEG_CHANNEL_SUBSET)
        file_id = [cfg.modality_file_id, f"Channels_{'_'.join(subset)}"]
        expected_ch = len(CMC_EEG_CHANNEL_SUBSET)
    else:
        file_id, expected_ch = cfg.modality_file_id, None
    spectrogram, times, freqs = fetch_stored_spectrograms(feat_dir, modality=cfg.modality, file_identifier=file_id,
                                                          expected_n_channels=expected_ch)
    times_arr = np.asarray(times, dtype=np.float64)
    if cfg.use_stretched_window_timestamps:
        half = 0.5 * (cfg.cmc_time_window_sec if cfg.modality == "CMC" else cfg.psd_time_window_sec)
        timestamps = xlog.add_time_index(start_timestamp=qtc_start + pd.Timedelta(seconds=half),
                                         end_timestamp=qtc_end - pd.Timedelta(seconds=half),
                                         n_timesteps=len(times_arr))
    else:
        # window centres are stored as seconds since the measurement start; non-finite centres (outside-task
        # slots) become NaT and never match a trial span
        timestamps = pd.DatetimeIndex([qtc_start + pd.Timedelta(seconds=float(sec)) if np.isfinite(sec) else pd.NaT
                                       for sec in times_arr])
    return spectrogram, freqs, xlog.make_timezone_aware(timestamps), log_df


def _get_trial_spans(log_df) -> dict:
    """{trial_id: (start, end)} with the default latency / transient cut-off (cbpa.py:354-361)."""
    from . import experiment_log as xlog
    return xlog.get_all_task_start_ends(log_df, "dict")


def _common_time_grid_from_spans(cfg: CBPAConfig, trial_spans: dict, overlap_ratio=.5) -> np.ndarray:
    """Within-trial time grid at the spectrogram step, from the first trial's duration (cbpa.py:364-378)."""
    import pandas as pd
    tw = cfg.psd_time_window_sec if cfg.modality == "PSD" else cfg.cmc_time_window_sec
    first_start, first_end = next(iter(trial_spans.values()))
    dur = (pd.Timestamp(first_end) - pd.Timestamp(first_start)).total_seconds()
    n_times = max(1, int(dur / (tw * overlap_ratio)))
    return np.arange(n_times) * (tw * overlap_ratio)


def _band_power_per_trial(cfg: CBPAConfig, band_power: np.ndarray, timestamps, trial_spans: dict,
                          target_n_times: int | None):
    """(n_trials, n_times, n_channels) slices at native resolution, trials of another length linearly resampled
    onto ``target_n_times`` points (default: the modal length) - cbpa.py:381-432."""
    import warnings
    import pandas as pd
    slices, ids = [], []
    for trial_id, (t_start, t_end) in trial_spans.items():
        slc = band_power[np.asarray((timestamps >= t_start) & (timestamps < t_end))]
        if slc.shape[0] == 0:
            warnings.warn(f"  [WARN] Trial {trial_id}: no spectrogram windows found in span. Skipping.")
            continue
        slices.append(slc)
        ids.append(trial_id)
    if not slices:
        raise RuntimeError("No trial windows found — check timestamp alignment.")
    if target_n_times is None:
        target_n_times = int(pd.Series([x.shape[0] for x in slices]).mode().iloc[0])
    n_ch = slices[0].shape[-1]
    result = np.full((len(slices), target_n_times, n_ch), np.nan)
    dst_x = np.linspace(0, 1, target_n_times)
    for i, slc in enumerate(slices):
        if slc.shape[0] == target_n_times:
            result[i] = slc
        else:
            src_x = np.linspace(0, 1, slc.shape[0])
            for ch in range(n_ch):
                result[i, :, ch] = np.interp(dst_x, src_x, slc[:, ch])
    return result, ids


STATS_FRAME_SEG_SUFFIX: str = "1seg"


def load_stats_frame(data_root):
    """Newest ``Combined Statistics 1seg`` CSV - the single source of trial-level condition labels (cbpa.py:445-492)."""
    import pandas as pd
    feature_dir = Path(data_root) / "data" / "precomputed_features"
    try:
        csv_path = filemgmt.most_recent_file(feature_dir, ".csv", [f"Combined Statistics {STATS_FRAME_SEG_SUFFIX}"])
    except (ValueError, FileNotFoundError):
        raise FileNotFoundError(
            f"\n[CBPA] Required statistics frame not found in:\n  {feature_dir}\n"
            f"Expected a file matching 'Combined Statistics {STATS_FRAME_SEG_SUFFIX}' with extension '.csv'.\n"
            f"Please run the main statistical_workflow.py pipeline first (with n_within_trial_segments=1) to "
            f"generate it.")
    df = pd.read_csv(csv_path)
    missing = {"Subject ID", "Trial ID", "Category or Silence", "Perceived Category", "Music Listening"} - set(df.columns)
    if missing:
        raise ValueError(f"[CBPA] Statistics frame is missing required columns: {missing}\n  Loaded from: {csv_path}")
    print(f"  [stats frame] Loaded: {Path(csv_path).name}  ({len(df)} rows, {df['Subject ID'].nunique()} subjects, "
          f"{df['Trial ID'].nunique()} unique trial IDs)")
    return df


def get_trial_condition_map(stats_df, subject_id: int, condition_column: str) -> dict:
    """Trial ID -> condition label (None for NaN) of one subject, cbpa.py:495-529."""
    import pandas as pd
    subj = stats_df[stats_df["Subject ID"] == subject_id]
    if subj.empty:
        raise ValueError(f"[CBPA] Subject {subject_id} not found in statistics frame. "
                         f"Available subjects: {sorted(stats_df['Subject ID'].unique())}")
    out = {}
    for _, row in subj.iterrows():
        val = row.get(condition_column, None)
        out[int(row["Trial ID"])] = None if pd.isna(val) else str(val)
    return out


# ----------------------------------------------------------------------------- runner
def build_contrast_array(cfg: CBPAConfig):
    """Per-subject A - B contrast ``X (n_subjects, n_times, n_channels)``, channel names and the time (seconds) or
    phase (degrees) grid from the stored spectrogram files, the enriched logs and the Combined Statistics frame -
    cbpa.py:733-942.  Subjects whose files fail to load, who are missing from the statistics frame or lack one of
    the two conditions are skipped with a warning (which lowers the degrees of freedom of the threshold)."""
    import warnings
    stats_df = load_stats_frame(cfg.data_root)
    subjects = sorted(stats_df["Subject ID"].astype(int).unique())
    if cfg.exclude_subjects:
        print(f"  [Exclusions] Skipping subjects: {cfg.exclude_subjects}")
        subjects = [s for s in subjects if s not in cfg.exclude_subjects]
    print(f"  [subjects] Running on {len(subjects)} subjects: {subjects}")
    if cfg.modality == "CMC":
        # stored CMC spectrograms already hold the motor subset: never index them with 64-channel indices
        ch_indices = None
        ch_names_out = cfg.channels if cfg.channels is not None else CMC_EEG_CHANNEL_SUBSET
    elif cfg.channels is not None:
        ch_indices, ch_names_out = [EEG_CHANNEL_IND_DICT[ch] for ch in cfg.channels], cfg.channels
    else:
        ch_indices, ch_names_out = None, EEG_CHANNELS
    time_grid, n_times_ref = None, None
    if cfg.use_phase_normalization:
        time_grid = np.linspace(0, 360, cfg.n_phase_bins, endpoint=False)
        n_times_ref = cfg.n_phase_bins
    diffs = []
    for subj in subjects:
        print(f"  Subject {subj:02} — loading...")
        try:
            spectrogram, freqs, timestamps, log_df = _load_subject_data(cfg, subj)
        except Exception as exc:
            warnings.warn(f"Subject {subj:02}: load failed ({exc}). Skipping.")
            continue
        try:
            cond_map = get_trial_condition_map(stats_df, subj, cfg.condition_column)
        except ValueError as exc:
            warnings.warn(str(exc) + " Skipping.")
            continue
        spans = {int(k): v for k, v in _get_trial_spans(log_df).items()}
        missing = set(spans) - set(cond_map)
        if missing:
            warnings.warn(f"Subject {subj:02}: trial IDs {missing} present in log_df but missing from stats frame. "
                          f"These trials will be skipped.")
        if time_grid is None:
            time_grid = _common_time_grid_from_spans(cfg, spans, overlap_ratio=cfg.overlap_ratio)
            n_times_ref = len(time_grid)
        band_power = _extract_band_power(cfg, spectrogram, freqs, ch_indices)
        if cfg.use_phase_normalization:
            cycles = _band_power_per_phase(cfg, band_power, timestamps, spans, cond_map, log_df,
                                           min_cycle_coverage_ratio=.8)
            cyc_a, cyc_b = cycles.get(cfg.condition_A, []), cycles.get(cfg.condition_B, [])
            short = [(c, len(x)) for c, x in ((cfg.condition_A, cyc_a), (cfg.condition_B, cyc_b))
                     if len(x) < cfg.min_cycles_per_condition]
            if short:
                warnings.warn(f"Subject {subj:02}: only {short[0][1]} valid cycles for '{short[0][0]}' "
                              f"(min={cfg.min_cycles_per_condition}). Skipping.")
                continue
            diffs.append(phase_contrast_from_cycles(cfg, cycles))
            print(f"    → {len(cyc_a)} cycles '{cfg.condition_A}', {len(cyc_b)} cycles '{cfg.condition_B}'")
            continue
        trial_data, trial_ids = _band_power_per_trial(cfg, band_power, timestamps, spans, n_times_ref)
        idx_a = [i for i, t in enumerate(trial_ids) if cond_map.get(t) == cfg.condition_A]
        idx_b = [i for i, t in enumerate(trial_ids) if cond_map.get(t) == cfg.condition_B]
        empty = [c for c, idx in ((cfg.condition_A, idx_a), (cfg.condition_B, idx_b)) if len(idx) == 0]
        if empty:
            warnings.warn(f"Subject {subj:02}: no trials found for '{empty[0]}' in '{cfg.condition_column}'. Skipping.")
            continue
        diffs.append(np.nanmean(trial_data[idx_a], axis=0) - np.nanmean(trial_data[idx_b], axis=0))
        print(f"    → {len(idx_a)} trials '{cfg.condition_A}', {len(idx_b)} trials '{cfg.condition_B}'")
    if not diffs:
        raise RuntimeError("[CBPA] No valid subjects produced a contrast. "
                           "Check data paths, subject IDs, and condition labels.")
    X = np.stack(diffs, axis=0)
    print(f"\n  Contrast array built: {X.shape}  "
          f"[{X.shape[0]} subjects × {X.shape[1]} time pts × {X.shape[2]} channels]")
    return X, ch_names_out, time_grid


def default_spatial_adjacency(ch_names):
    """Stand-in for ``find_ch_adjacency(_build_mne_info(ch_names))`` (cbpa.py:200-243) without MNE: Delaunay
    neighbours of the cap layout ``channel_layout.EEG_POSITIONS`` (reference ``visualizations.py:61-131``) restricted
    to ``ch_names``.  MNE triangulates the standard_1020 montage instead, so the edge set can differ in detail;
    pass ``spatial_adjacency=`` to ``run_cbpa`` to use an exact one."""
    from .channel_layout import EEG_POSITIONS
    unknown = [c for c in ch_names if c not in EEG_POSITIONS]
    if unknown:
        raise ValueError(f"no sensor position for channels {unknown}; pass spatial_adjacency=")
    return find_ch_adjacency_from_positions(np.array([EEG_POSITIONS[c] for c in ch_names], dtype=np.float64))


def run_cbpa(cfg: CBPAConfig, cluster_rows_accumulator: list[dict] | None = None, *, contrast=None,
             spatial_adjacency=None, signs: np.ndarray | None = None) -> dict:
    """Full CBPA for one contrast, cbpa.py:985-1067: ``run_cbpa(cfg, cluster_rows_accumulator)`` as the reference
    calls it.  Returns the reference's result dict (keys t_obs, t_thresh, clusters, cluster_pv, H0,
    good_cluster_inds, ch_names, time_grid, cfg, n_valid_subjects).  Keyword-only extras (defaults reproduce the
    reference's behaviour): ``contrast=(X, ch_names, time_grid)`` skips the file loading, ``spatial_adjacency``
    (matrix or (n_ch, 2) positions) replaces the layout-derived channel adjacency, ``signs`` fixes the sign table."""
    filemgmt.assert_dir(cfg.output_dir)
    _print_header(cfg)
    X, ch_names, time_grid = contrast if contrast is not None else build_contrast_array(cfg)
    X = np.asarray(X, dtype=np.float64)
    n_subj, n_times, n_ch = X.shape
    df_stat = n_subj - 1
    if cfg.tail == 0:
        t_thresh = t_dist.ppf(1.0 - cfg.alpha_cluster_forming / 2, df=df_stat)
    else:
        t_thresh = t_dist.ppf(1.0 - cfg.alpha_cluster_forming, df=df_stat)
    print(f"\n  Cluster-forming threshold  t({df_stat}) = ±{t_thresh:.4f}  "
          f"(α = {cfg.alpha_cluster_forming}, tail = {cfg.tail})")
    rng = np.random.default_rng(cfg.seed)
    if spatial_adjacency is None:
        spatial_adjacency = default_spatial_adjacency(list(ch_names))
    adjacency = _build_adjacency(spatial_adjacency, n_times)
    if cfg.use_phase_normalization:
        adjacency = _add_phase_wraparound(adjacency, n_times, n_ch, np.asarray(time_grid))
    if cfg.use_spatio_temporal:
        t_obs, clusters, cluster_pv, H0 = spatio_temporal_cluster_1samp_test(
            X, n_permutations=cfg.n_permutations, threshold=t_thresh, tail=cfg.tail, adjacency=adjacency,
            n_jobs=cfg.n_jobs, seed=rng, out_type="mask", verbose=True, signs=signs)
    else:
        X_flat = X.reshape(n_subj, n_times * n_ch)
        print(f"  [adjacency] flat: {adjacency.shape}, nnz edges: {adjacency.nnz}")
        t_obs_flat, clusters, cluster_pv, H0 = permutation_cluster_1samp_test(
            X_flat, n_permutations=cfg.n_permutations, threshold=t_thresh, tail=cfg.tail,
            adjacency=adjacency, n_jobs=cfg.n_jobs, seed=rng, out_type="mask", verbose=True, signs=signs)
        t_obs = t_obs_flat.reshape(n_times, n_ch)
    alpha_cbpa = 0.05
    good_cluster_inds = np.where(np.array(cluster_pv) < alpha_cbpa)[0]
    print(f"\n  Clusters found: {len(clusters)} total, "
          f"{len(good_cluster_inds)} significant (cluster p < {alpha_cbpa})")
    for idx in good_cluster_inds:
        print(f"    Cluster #{idx + 1:02d}:  p = {cluster_pv[idx]:.4f}")
    results = dict(t_obs=t_obs, t_thresh=t_thresh, clusters=clusters, cluster_pv=np.array(cluster_pv), H0=H0,
                   good_cluster_inds=good_cluster_inds, ch_names=ch_names, time_grid=time_grid, cfg=cfg,
                   n_valid_subjects=n_subj)
    _save_results(results, cfg, cluster_rows_accumulator=cluster_rows_accumulator,
                  save_per_run_cluster_csv=(cluster_rows_accumulator is None))
    if cfg.save_plots or cfg.show_plots:
        _plot_results(results, cfg)
    return results


def _plot_results(results: dict, cfg: CBPAConfig) -> None:
    """cbpa.py:1064-1065 hands the result dict to ``visualizations.plot_cbpa_results`` (matplotlib, out of scope of
    this package): use the reference's plotting module when it is importable, otherwise say so once and go on."""
    try:
        import importlib
        viz = importlib.import_module("src.pipeline.visualizations")
        viz.plot_cbpa_results(results, cfg)
    except Exception as exc:                                       # no reference tree / no matplotlib backend
        print(f"  [plots] skipped ({type(exc).__name__}): plotting lives in the reference's visualizations module")


def _cluster_mask(cluster, n_times: int, n_ch: int) -> np.ndarray:
    if isinstance(cluster, np.ndarray) and cluster.dtype == bool:
        return cluster.reshape(n_times, n_ch) if cluster.ndim == 1 else cluster
    mask = np.zeros((n_times, n_ch), dtype=bool)
    mask[cluster] = True
    return mask


def _save_results(results: dict, cfg: CBPAConfig, cluster_rows_accumulator: list[dict] | None = None,
                  save_per_run_cluster_csv: bool = False) -> None:
    """.npz archive, t_obs CSV and cluster summary rows with the reference's column names
    (cbpa.py:1076-1185) so ``statistical_reporting._section_cbpa`` keeps reading them."""
    import pandas as pd
    stem = filemgmt.file_title(cfg.hypothesis_label, "")
    out_dir = Path(cfg.output_dir)
    np.savez(out_dir / (stem + ".npz"), t_obs=results["t_obs"], cluster_pv=results["cluster_pv"],
             H0=results["H0"], ch_names=results["ch_names"], time_grid=results["time_grid"],
             good_cluster_inds=results["good_cluster_inds"])
    t_obs, time_grid, ch_names = results["t_obs"], results["time_grid"], list(results["ch_names"])
    t_ax = np.asarray(time_grid) if time_grid is not None else np.arange(t_obs.shape[0])
    pd.DataFrame(t_obs, index=pd.Index(np.round(t_ax, 4), name="time_s"), columns=ch_names).to_csv(
        out_dir / (stem + "_t_obs.csv"))
    n_times, n_ch = t_obs.shape
    axis_label = "phase_deg" if cfg.use_phase_normalization else "time_s"
    rows = []
    for idx, (cluster, pv) in enumerate(zip(results["clusters"], results["cluster_pv"])):
        mask = _cluster_mask(cluster, n_times, n_ch)
        t_in = np.where(mask.any(axis=1))[0]
        ch_in = np.where(mask.any(axis=0))[0]
        rows.append({
            "hypothesis": cfg.hypothesis_label, "modality": cfg.modality, "freq_band": cfg.freq_band,
            "condition_column": cfg.condition_column, "condition_A": cfg.condition_A,
            "condition_B": cfg.condition_B, "n_within_trial_segs": cfg.n_within_trial_segs,
            "n_permutations": cfg.n_permutations, "alpha_cluster_forming": cfg.alpha_cluster_forming,
            "tail": cfg.tail, "n_valid_subjects": results["n_valid_subjects"],
            "cluster_index": idx + 1, "p_value": round(float(pv), 6),
            "significant": bool(idx in results["good_cluster_inds"]),
            "peak_t": round(float(np.abs(t_obs[mask]).max()) if mask.any() else 0.0, 4),
            "t_thresh": round(float(results["t_thresh"]), 4), "n_time_points": int(len(t_in)),
            f"{axis_label}_start": round(float(t_ax[t_in[0]]), 4) if len(t_in) > 0 else None,
            f"{axis_label}_end": round(float(t_ax[t_in[-1]]), 4) if len(t_in) > 0 else None,
            "n_channels": int(len(ch_in)), "channels": "; ".join(ch_names[i] for i in ch_in),
        })
    if cluster_rows_accumulator is not None:
        cluster_rows_accumulator.extend(rows)
    if save_per_run_cluster_csv:
        pd.DataFrame(rows).to_csv(out_dir / (stem + "_cluster_summary.csv"), index=False)


def _print_header(cfg: CBPAConfig) -> None:
    bar = "═" * 70
    print(f"\n{bar}\n  CBPA: {cfg.hypothesis_label}\n  Feature    : {cfg.modality} | {cfg.freq_band} band\n"
          f"  Contrast   : '{cfg.condition_A}'  −  '{cfg.condition_B}'\n"
          f"  Permutations: {cfg.n_permutations}  |  tail={cfg.tail}  |  α={cfg.alpha_cluster_forming}\n{bar}\n")


def run_batch(configs: list[CBPAConfig], **run_kwargs):
    """Sequential runs + one combined cluster-summary CSV, cbpa.py:1214-1250."""
    import pandas as pd
    all_results, rows = [], []
    for i, cfg in enumerate(configs):
        print(f"\n[{i + 1}/{len(configs)}] Starting: {cfg.hypothesis_label}")
        all_results.append(run_cbpa(cfg, cluster_rows_accumulator=rows, **run_kwargs))
    combined = pd.DataFrame(rows)
    if not combined.empty:
        out_path = Path(configs[0].output_dir) / filemgmt.file_title("CBPA Combined Cluster Summary", ".csv")
        combined.to_csv(out_path, index=False)
    return all_results, combined
