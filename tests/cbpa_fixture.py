"""Deterministic synthetic study directory in the reference's on-disk layout (what ``cbpa.build_contrast_array``,
reference ``src/pipeline/cbpa.py:733-942``, reads): stored spectrogram ``.npy`` trios, enriched experiment logs,
subject JSONs and the Combined Statistics frame.  Shared by ``scripts/make_golden.py`` (which runs the UNMODIFIED
reference on it to produce ``tests/golden/contrast.npz``) and by the tests (which run this repo on the same files).
Only file-name time stamps differ between two builds, never contents."""
from __future__ import annotations

import json
from datetime import datetime
from pathlib import Path

import numpy as np
import pandas as pd

CMC_SUBSET = ["C5", "C3", "C1", "FC5", "FC3", "FC1", "F3", "CP5", "CP3", "CP1", "P3"]
CMC_SUBSET_MIRRORED = ["C6", "C4", "C2", "FC6", "FC4", "FC2", "F4", "CP6", "CP4", "CP2", "P4"]
T0 = pd.Timestamp("2025-03-04 10:00:00")
N_TRIALS, TRIAL_SEC, GAP_SEC, LEAD_SEC = 8, 40, 6, 12
CATEGORIES = ["Happy", "Silence", "Sad", "Silence", "Happy", "Silence", "Sad", "Silence"]


def _stamp(k: int = 0) -> str:
    return datetime(2025, 3, 5, 9, 0, k).strftime("%Y-%m-%d %H_%M_%S")


def _log_frame(subject: int, excluded_trial: int | None, drop_condition: str | None) -> pd.DataFrame:
    """One row per second; triggers 2 s in / 2 s before the end; music trials carry a Song ID, silence trials a
    Silence ID; 'Task Frequency' is set while the motor task runs (both kinds)."""
    total = LEAD_SEC + N_TRIALS * (TRIAL_SEC + GAP_SEC) + 10
    rows = []
    song = silence = 0
    ids = {}
    for k, cat in enumerate(CATEGORIES):
        if cat == "Silence":
            ids[k] = (np.nan, silence)
            silence += 1
        else:
            ids[k] = (song, np.nan)
            song += 1
    for sec in range(total):
        row = {"Time": T0 + pd.Timedelta(seconds=sec), "Event": np.nan, "Trial ID": np.nan, "Song ID": np.nan,
               "Silence ID": np.nan, "Song Skipped": False, "Trial Exclusion Bool": False, "Task Frequency": np.nan,
               "Song Title": np.nan}
        if sec == 2:
            row["Event"] = "Start Trigger"
        if sec == total - 2:
            row["Event"] = "Stop Trigger"
        rel = sec - LEAD_SEC
        if rel >= 0:
            k, inside = divmod(rel, TRIAL_SEC + GAP_SEC)
            if k < N_TRIALS and inside < TRIAL_SEC and not (drop_condition and CATEGORIES[k] == drop_condition):
                row["Trial ID"] = k
                row["Song ID"], row["Silence ID"] = ids[k]
                row["Song Title"] = f"song {k}" if CATEGORIES[k] != "Silence" else np.nan
                row["Task Frequency"] = 0.1 if inside >= 2 else np.nan      # task starts 2 s into the music
                if CATEGORIES[k] == "Silence":
                    row["Task Frequency"] = 0.1
                row["Trial Exclusion Bool"] = (k == excluded_trial)
        rows.append(row)
    return pd.DataFrame(rows)


def build_study(root, n_subjects: int = 6, seed: int = 123) -> Path:
    """Writes the study below ``root`` and returns it (pass it as ``CBPAConfig.data_root``).  Subject 2 is
    left-handed (mirrored CMC file names), subject 3 has trial 4 excluded, subject 5 never did a 'Happy' trial
    (skipped by a Happy - Silence contrast), subject 6 has no spectrogram files at all (load fails, skipped)."""
    root = Path(root)
    rng = np.random.default_rng(seed)
    feat = root / "data" / "precomputed_features"
    feat.mkdir(parents=True, exist_ok=True)
    stats_rows = []
    freqs_cmc = np.arange(0.0, 60.5, 0.5)
    freqs_psd = np.arange(0.0, 64.0, 4.0)
    for subj in range(1, n_subjects + 1):
        exp = root / "data" / "experiment_results" / f"subject_{subj:02}"
        (exp / "experiment_logs").mkdir(parents=True, exist_ok=True)
        sfeat = feat / f"subject_{subj:02}"
        sfeat.mkdir(parents=True, exist_ok=True)
        left = subj == 2
        log = _log_frame(subj, excluded_trial=4 if subj == 3 else None, drop_condition="Happy" if subj == 5 else None)
        log.to_csv(exp / "experiment_logs" / f"{_stamp()} Enriched Experiment Log.csv", index=False)
        with open(exp / f"{_stamp()} Subject {subj:02} Data.json", "w") as fh:
            json.dump({"Name": "x", "Birthdate": "y", "Gender": "d", "Dominant hand": "Left" if left else "Right",
                       "Listening habit": "Seldom", "Dancing habit": 1, "Athleticism": 2}, fh)
        with open(exp / f"{_stamp()} Post-Study Feedback Data.json", "w") as fh:
            json.dump({"Total fatigue": 2, "Total pleasure": 5}, fh)
        for k, cat in enumerate(CATEGORIES):
            if subj == 5 and cat == "Happy":
                continue
            stats_rows.append({"Subject ID": subj, "Trial ID": k, "Category or Silence": cat,
                               "Perceived Category": np.nan if cat == "Silence" else cat,
                               "Music Listening": cat != "Silence", "Segment ID": 0})
        total = len(log)
        # CMC: 2 s windows, 1 s step -> centres 1, 2, ... s after the Start Trigger + 0.75 s latency
        n_w = total - 6
        centers = 1.0 + np.arange(n_w, dtype=np.float64)
        centers[5] = np.nan                                            # an outside-task slot
        spec = rng.random((n_w, len(freqs_cmc), 11)).astype(np.float32) * 0.3
        happy = np.zeros(n_w)
        for k, cat in enumerate(CATEGORIES):
            if cat == "Happy":
                a = LEAD_SEC + k * (TRIAL_SEC + GAP_SEC)
                happy[a: a + TRIAL_SEC] = 1.0
        spec[:, 26:60, :4] += 0.4 * happy[:, None, None] * np.sin(2 * np.pi * 0.1 * centers[:, None, None]) ** 2
        np.nan_to_num(spec, copy=False)
        subset = CMC_SUBSET_MIRRORED if left else CMC_SUBSET
        suffix = f"Channels_{'_'.join(subset)}"
        if subj != 6:
            _save_trio(sfeat, spec, centers, freqs_cmc, "Flexor CMC Trial-wise", suffix, 1.0, k0=1)
            # PSD: 0.25 s windows, 0.125 s step, 64 channels, log-scaled
            n_p = (total - 6) * 8
            c_p = 0.125 + 0.125 * np.arange(n_p, dtype=np.float64)
            psd = rng.normal(-1.0, 0.2, (n_p, len(freqs_psd), 64)).astype(np.float32)
            psd[:, 2:4, 20:30] += 0.5 * np.repeat(happy, 8)[:n_p, None, None]
            _save_trio(sfeat, psd, c_p, freqs_psd, "eeg PSD", "All_Channels", 0.125, k0=4)
    pd.DataFrame(stats_rows).to_csv(feat / f"{_stamp()} Combined Statistics 1seg.csv", index=False)
    return root


def _save_trio(folder: Path, spec, centers, freqs, modality: str, suffix: str, step: float, k0: int) -> None:
    """The three files of ``signal_features.save_spectrograms`` (reference ``:1033-1046``) with fixed time stamps."""
    np.save(folder / f"{_stamp(k0)} {modality} Spectrograms {spec.shape[2]}ch {step:.2f}sec_step {suffix}.npy", spec)
    np.save(folder / f"{_stamp(k0 + 1)} {modality} Timecenters {len(centers)}windows {suffix}.npy", centers)
    np.save(folder / f"{_stamp(k0 + 2)} {modality} Frequencies {len(freqs)}freqs {suffix}.npy", freqs)


def configs(cbpa_module, root, out_dir):
    """The three contrasts of the golden file, as ``CBPAConfig`` objects of the given module (this repo's or the
    reference's)."""
    common = dict(condition_column="Category or Silence", condition_A="Happy", condition_B="Silence",
                  data_root=Path(root), output_dir=Path(out_dir), save_plots=False, n_permutations=64, seed=42)
    return {
        "cmc_clock": cbpa_module.CBPAConfig(modality="CMC", modality_file_id="Flexor", freq_band="beta",
                                            hypothesis_label="cmc clock", **common),
        "cmc_phase": cbpa_module.CBPAConfig(modality="CMC", modality_file_id="Flexor", freq_band="beta",
                                            use_phase_normalization=True, n_phase_bins=12, min_cycles_per_condition=2,
                                            hypothesis_label="cmc phase", **common),
        "psd_alpha": cbpa_module.CBPAConfig(modality="PSD", modality_file_id="eeg", freq_band="alpha",
                                            channels=["C3", "C1", "Cz", "C2", "C4", "CP1", "CPz", "CP2"],
                                            hypothesis_label="psd alpha", **common),
    }
