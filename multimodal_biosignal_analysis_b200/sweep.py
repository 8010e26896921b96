"""The full cortico-muscular coherence sweep of BASELINE config 5 as ONE public call: per subject-condition the
all-pairs Welch coherence with its surrogate null, then the group-level cluster-based permutation analysis of the
condition contrast.

Reference shape of the work: the subject loop of ``src/subject_feature_extraction_workflow.py:37`` with the per-subject
CMC call at ``:246-255`` (feature extraction), the EMG-max reduction of ``signal_features.py:1132-1171`` and the
contrast + CBPA of ``src/pipeline/cbpa.py:858-879, 985-1067`` (statistics) - there two scripts with ``.npy`` files in
between, here one pass that never leaves the GPUs:

  units (subject, condition) are dealt round-robin over the ranks, NO collective while they run;
  per unit: upload -> K1 x 2 -> K2 -> operand planes -> the whole surrogate null -> download of coherence / counts /
  maxima, with the upload of unit i + 1 and the download of unit i - 1 overlapped (``data_surrogation.surrogate_null_sweep``);
  the per-unit (F, Ne) EMG-max maps (2 MB for 80 units) are summed into every rank with one all-reduce;
  the A - B contrast (n_subjects, F, Ne) goes through the CBPA kernels with the permutations sharded over the ranks
  and one all-gather of the H0 slices (``cbpa.permutation_cluster_1samp_test``).
"""
from __future__ import annotations

import numpy as np
import torch
from scipy.stats import t as t_dist

from . import cbpa as cb
from . import data_surrogation as dsur
from . import dist as cdist
from .channel_layout import EEG_CHANNELS


def cmc_surrogate_cbpa_sweep(units, sampling_freq: float, nperseg: int = 2048, noverlap: int | None = None,
                             freq_band: tuple[float, float] | None = (1.0, 100.0), segment_starts=None,
                             n_surrogates: int = 10000, mode: str = "phase", seed: int = 0, alpha: float = 0.05,
                             mask_nonsignificant: bool = False, contrasts=None, spatial_adjacency=None,
                             n_permutations: int = 10000, tail: int = 0, alpha_cluster_forming: float = 0.05,
                             cbpa_seed: int = 42, signs=None, keep_pair_results: bool = False) -> dict:
    """``units``: mapping ``(subject, condition) -> (eeg, emg)`` (or ``-> callable`` returning them, for lazy loading)
    of equally shaped time-first recordings.  Returns a dict with

      keys                 list of (subject, condition) in processing order
      freqs                (F,)
      cmc                  (n_units, F, Ne) float32: max over EMG channels of the coherence (zeroed where the surrogate
                           p-value is >= alpha when ``mask_nonsignificant``) - what the reference stores per subject
      threshold_fwe        (n_units,) family-wise surrogate threshold per unit
      n_significant_pairs  (n_units,) number of (f, i, j) with p < alpha
      cbpa                 {(cond_A, cond_B): dict(t_obs, clusters, cluster_pv, H0, t_thresh, subjects, X)} for every
                           requested contrast (default: the first two conditions)
      pairs                only with ``keep_pair_results``: {key: dict(coherence, p_values)} of the units THIS rank ran

    Every rank returns the same ``cmc`` / ``cbpa``; surrogates of unit u are seeded with ``seed + u`` and sign flips
    come from one host table, so the result does not depend on the number of ranks."""
    keys = list(units.keys())
    n_units = len(keys)
    if n_units == 0:
        raise ValueError("no units")
    rank, world = cdist.world()
    mine = list(cdist.round_robin(n_units, rank, world))

    def load(u):
        item = units[keys[u]]
        return item() if callable(item) else item

    # group-level inputs that do not depend on the data (channel adjacency -> lattice adjacency -> device CSR, sign
    # table) are prepared by a host thread while the unit stage keeps the GPU busy
    subjects = list(dict.fromkeys(k[0] for k in keys))
    conditions = list(dict.fromkeys(k[1] for k in keys))
    if contrasts is None:
        contrasts = [(conditions[0], conditions[1])] if len(conditions) >= 2 else []
    index = {k: u for u, k in enumerate(keys)}
    prep = _GroupPrep(contrasts, subjects, index, load(0), sampling_freq, nperseg, freq_band, spatial_adjacency,
                      n_permutations, tail, cbpa_seed, signs)

    cmc = None
    thr = np.zeros(n_units, dtype=np.float64)
    nsig = np.zeros(n_units, dtype=np.float64)
    freqs = None
    pairs = {}
    gen = dsur.surrogate_null_sweep((load(u) for u in mine), sampling_freq, nperseg=nperseg, noverlap=noverlap,
                                    freq_band=freq_band, segment_starts=segment_starts, n_surrogates=n_surrogates,
                                    mode=mode, seed=seed, alpha=alpha, unit_indices=mine)
    for res in gen:
        u = res["unit"]
        coh, p = res["coherence"], res["p_values"]
        if cmc is None:
            cmc = np.zeros((n_units,) + coh.shape[:2], dtype=np.float32)
            freqs = res["freqs"]
        sig = p < alpha
        cmc[u] = (np.where(sig, coh, np.float32(0)) if mask_nonsignificant else coh).max(axis=2)
        thr[u], nsig[u] = res["threshold_fwe"], float(sig.sum())
        if keep_pair_results:
            pairs[keys[u]] = dict(coherence=coh.copy(), p_values=p.copy())
    if world > 1:
        # the only exchange of the unit stage: every rank contributes its rows (all others are zero)
        first = load(0)
        shape = _map_shape(first, sampling_freq, nperseg, freq_band)
        if cmc is None:                                            # more ranks than units
            cmc = np.zeros((n_units,) + shape, dtype=np.float32)
        dev = torch.device("cuda", torch.cuda.current_device())
        packed = torch.from_numpy(np.concatenate([cmc.reshape(-1), thr.astype(np.float32),
                                                  nsig.astype(np.float32)])).to(dev)
        cdist.all_reduce_sum_(packed)
        h = packed.cpu().numpy()
        cmc = h[: cmc.size].reshape(cmc.shape).copy()
        thr = h[cmc.size: cmc.size + n_units].astype(np.float64)
        nsig = h[cmc.size + n_units:].astype(np.float64)
        if freqs is None:
            f = np.fft.rfftfreq(nperseg, d=1 / sampling_freq)
            freqs = f if freq_band is None else f[(f >= freq_band[0]) & (f <= freq_band[1])]

    # ---- group level: A - B contrast per subject -> CBPA over (frequency, EEG channel) ----
    out_cbpa = {}
    adjacency = prep.result()
    for cond_a, cond_b in contrasts:
        subj = prep.subjects_of[(cond_a, cond_b)]
        X = np.stack([cmc[index[(s, cond_a)]].astype(np.float64) - cmc[index[(s, cond_b)]].astype(np.float64)
                      for s in subj])
        q = alpha_cluster_forming / 2 if tail == 0 else alpha_cluster_forming
        t_thresh = float(t_dist.ppf(1.0 - q, df=len(subj) - 1)) * (-1.0 if tail == -1 else 1.0)
        t_obs, clusters, pv, H0 = cb.spatio_temporal_cluster_1samp_test(
            X, threshold=t_thresh, n_permutations=n_permutations, tail=tail, adjacency=adjacency, out_type="mask",
            signs=prep.signs[(cond_a, cond_b)])
        out_cbpa[(cond_a, cond_b)] = dict(t_obs=t_obs, clusters=clusters, cluster_pv=pv, H0=H0, t_thresh=t_thresh,
                                          subjects=subj, X=X)
    out = dict(keys=keys, freqs=freqs, cmc=cmc, threshold_fwe=thr, n_significant_pairs=nsig.astype(np.int64),
               cbpa=out_cbpa, n_surrogates=n_surrogates, n_permutations=n_permutations)
    if keep_pair_results:
        out["pairs"] = pairs
    return out


def _map_shape(first_item, sampling_freq, nperseg, freq_band):
    f = np.fft.rfftfreq(nperseg, d=1 / sampling_freq)
    n_f = len(f) if freq_band is None else int(((f >= freq_band[0]) & (f <= freq_band[1])).sum())
    return n_f, int(first_item[0].shape[1])


class _GroupPrep:
    """Everything the CBPA needs besides the contrast itself, built on a host thread: which subjects enter each
    contrast, the (frequency x channel) lattice adjacency as device CSR, and the sign tables (``cbpa.make_sign_table``
    with ``default_rng(cbpa_seed)``, exactly what ``permutation_cluster_1samp_test`` would draw)."""

    def __init__(self, contrasts, subjects, index, first_item, sampling_freq, nperseg, freq_band, spatial_adjacency,
                 n_permutations, tail, cbpa_seed, signs):
        import threading
        self.subjects_of, self.signs = {}, {}
        for cond_a, cond_b in contrasts:
            subj = [s for s in subjects if (s, cond_a) in index and (s, cond_b) in index]
            if len(subj) < 2:
                raise ValueError(f"contrast {cond_a!r} - {cond_b!r}: fewer than two subjects have both conditions")
            self.subjects_of[(cond_a, cond_b)] = subj
        n_f, n_e = _map_shape(first_item, sampling_freq, nperseg, freq_band)
        if contrasts and spatial_adjacency is None and n_e != len(EEG_CHANNELS):
            raise ValueError("spatial_adjacency is required unless the EEG array holds the 64 channels of the cap")
        self._adj, self._err = None, None
        device = torch.cuda.current_device()

        def work():
            try:
                torch.cuda.set_device(device)
                sp = spatial_adjacency if spatial_adjacency is not None else cb.default_spatial_adjacency(EEG_CHANNELS)
                self._adj = cb.DeviceAdjacency(cb.combine_adjacency(n_f, sp))
                for key, subj in self.subjects_of.items():
                    self.signs[key] = signs if signs is not None else cb.make_sign_table(
                        n_permutations, len(subj), np.random.default_rng(cbpa_seed), tail)
            except Exception as exc:                              # noqa: BLE001 - re-raised by result()
                self._err = exc

        self._thread = None
        if contrasts:
            self._thread = threading.Thread(target=work, name="cmc-group-prep", daemon=True)
            self._thread.start()

    def result(self):
        if self._thread is not None:
            self._thread.join()
        if self._err is not None:
            raise self._err
        return self._adj
