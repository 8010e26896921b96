"""The handful of experiment-log readers that ``cbpa._load_subject_data`` / ``build_contrast_array`` (reference
``src/pipeline/cbpa.py:282-361, 733-942``) need, restated compactly on pandas so that ``run_batch(CONTRASTS)`` runs
without the reference tree.  Same names, arguments, defaults and results as the reference functions cited below;
everything else of ``data_integration.py`` / ``data_analysis.py`` (raw-log integration, questionnaires, validation,
plotting helpers) stays out of scope - the inputs here are the already integrated per-subject files the reference's
``data_integration_workflow.py`` writes:

  data/experiment_results/subject_XX/experiment_logs/"<stamp> ... Enriched Experiment Log ... .csv"
  data/experiment_results/subject_XX/"<stamp> ... Subject ... Data ... .json", "... Post-Study Feedback Data ... .json"
"""
from __future__ import annotations

import json
from pathlib import Path
from typing import Literal

import numpy as np
import pandas as pd

from . import file_management as filemgmt


def make_timezone_aware(dt_index, timezone: str = "utc"):
    """Naive DatetimeIndex / Series-with-DatetimeIndex / Timestamp -> localised; aware input is returned unchanged
    (``data_analysis.py:686-800``)."""
    timezone = timezone.lower()
    if isinstance(dt_index, pd.DatetimeIndex):
        return dt_index if dt_index.tz is not None else dt_index.tz_localize(timezone)
    if isinstance(dt_index, pd.Series):
        if not isinstance(dt_index.index, pd.DatetimeIndex):
            raise TypeError(f"Series must have a DatetimeIndex, got {type(dt_index.index)}")
        if dt_index.index.tz is not None:
            return dt_index
        out = dt_index.copy()
        out.index = out.index.tz_localize(timezone)
        return out
    if isinstance(dt_index, pd.Timestamp):
        return dt_index if dt_index.tz is not None else dt_index.tz_localize(timezone)
    raise TypeError("dt_index must be pd.DatetimeIndex, pd.Series with DatetimeIndex, "
                    f"or pd.Timestamp, got {type(dt_index)}")


def add_time_index(start_timestamp: pd.Timestamp, end_timestamp: pd.Timestamp, target_array=None,
                   n_timesteps: int | None = None):
    """Evenly spaced DatetimeIndex over [start, end] (``data_analysis.py:451-683``); with ``target_array`` the data
    come back as a Series / DataFrame on that index."""
    for name, ts in (("start_timestamp", start_timestamp), ("end_timestamp", end_timestamp)):
        if not isinstance(ts, pd.Timestamp):
            raise TypeError(f"{name} must be pd.Timestamp, got {type(ts)}")
    if start_timestamp >= end_timestamp:
        raise ValueError(f"start_timestamp ({start_timestamp}) must be strictly before end_timestamp ({end_timestamp})")
    if (start_timestamp.tz is None) != (end_timestamp.tz is None):
        raise ValueError("start_timestamp and end_timestamp must have matching timezone awareness: "
                         f"start_timestamp.tz={start_timestamp.tz}, end_timestamp.tz={end_timestamp.tz}")
    if target_array is not None:
        if isinstance(target_array, np.ndarray) and target_array.ndim != 1:
            raise ValueError(f"target_array must be 1-dimensional, got shape {target_array.shape}")
        if not isinstance(target_array, (pd.Series, pd.DataFrame, np.ndarray)):
            raise TypeError(f"target_array must be pd.Series, pd.DataFrame, or np.ndarray, got {type(target_array)}")
        if len(target_array) == 0:
            raise ValueError("target_array cannot be empty")
        n_timesteps = len(target_array)
    else:
        if n_timesteps is None:
            raise ValueError("Either target_array or n_timesteps must be provided. "
                             "If target_array is None, n_timesteps must be a positive integer.")
        if not isinstance(n_timesteps, (int, np.integer)):
            raise TypeError(f"n_timesteps must be an integer, got {type(n_timesteps)} with value {n_timesteps}")
        if n_timesteps <= 0:
            raise ValueError(f"n_timesteps must be a positive integer, got {n_timesteps}")
    index = pd.date_range(start=start_timestamp, end=end_timestamp, periods=n_timesteps)
    if target_array is None:
        return index
    if isinstance(target_array, pd.DataFrame):
        out = target_array.copy()
        out.index = index
        return out
    return pd.Series(np.asarray(target_array), index=index)


def fetch_personal_data(experiment_data_dir, include_name_and_birthdate: bool = False) -> dict:
    """Subject-data + post-study-feedback JSONs merged into one dict (``data_integration.py:1975-2003``);
    the CBPA loader reads ``'Dominant hand'`` from it."""
    experiment_data_dir = Path(experiment_data_dir)
    with open(filemgmt.most_recent_file(experiment_data_dir, ".json", ["Subject", "Data"])) as fh:
        data = json.load(fh)
    if not include_name_and_birthdate:
        data.pop("Name")
        data.pop("Birthdate")
    with open(filemgmt.most_recent_file(experiment_data_dir, ".json", ["Post-Study Feedback Data"])) as fh:
        data.update(json.load(fh))
    data.setdefault("Musical skill", 0)
    data["Listening habit [0-3]"] = {"Most of the day": 3, "A small part of the day": 2, "Every 2 or 3 days": 1,
                                     "Seldom": 0}[data["Listening habit"]]
    return data


def fetch_enriched_log_frame(experiment_data_dir, set_time_index: bool = True, verbose: bool = True) -> pd.DataFrame:
    """Newest ``Enriched Experiment Log`` CSV of the subject, ``Time`` as timezone-aware index
    (``data_integration.py:2006-2089``; the verbose trial listing of the reference is reduced to one line)."""
    log_dir = Path(experiment_data_dir) / "experiment_logs"
    try:
        log_frame = pd.read_csv(filemgmt.most_recent_file(log_dir, ".csv", ["Enriched Experiment Log"]))
    except ValueError:
        raise ValueError("Couldn't find enriched (integrated) experiment log frame with signature 'Enriched Experiment "
                         f"Log' in file title within {log_dir}...\nPlease ensure to run data_integration_workflow.py "
                         "on subject data beforehand.")
    if set_time_index:
        log_frame["Time"] = pd.to_datetime(log_frame["Time"])
        log_frame = log_frame.set_index("Time")
        log_frame.index = make_timezone_aware(log_frame.index)
    if verbose:
        print(f"Imported enriched log frame from {experiment_data_dir} ({len(log_frame)} rows, "
              f"{int(log_frame['Trial ID'].max() + 1)} trials)")
    return log_frame


def _as_utc(ts: pd.Timestamp) -> pd.Timestamp:
    return ts.tz_localize("UTC") if ts.tz is None else ts.tz_convert("UTC")


def get_qtc_measurement_start_end(df: pd.DataFrame, verbose: bool = True, assumed_latency_sec: float = .75):
    """(start, end) of the EEG / EMG measurement from the ``Start Trigger`` / ``Stop Trigger`` events (+ latency), an
    ``Actual Start Trigger`` overrides the start without latency; missing triggers fall back to the frame's first /
    last timestamp; always UTC (``data_integration.py:766-955``)."""
    if "Event" not in df.columns:
        raise KeyError("DataFrame must contain an 'Event' column with trigger information.")
    if not isinstance(df.index, pd.DatetimeIndex):
        if "Time" not in df.columns:
            raise ValueError('DataFrame must contain "Time" column or have a DatetimeIndex!')
        df["Time"] = pd.to_datetime(df["Time"])
        df.set_index("Time", inplace=True)

    def trigger(name, expected):
        hits = df.index[df["Event"] == name]
        if len(hits) > 1:
            raise ValueError(f"Found {len(hits)} '{name}' events. Expected {expected}.")
        return hits[0] if len(hits) else None

    latency = pd.Timedelta(seconds=assumed_latency_sec) if assumed_latency_sec > 0 else pd.Timedelta(0)
    start, stop = trigger("Start Trigger", "exactly one"), trigger("Stop Trigger", "exactly one")
    if start is None and verbose:
        print("No 'Start Trigger' event found, assuming measurement started at beginning")
    if stop is None and verbose:
        print("No 'Stop Trigger' event found, assuming measurement ran until end.")
    qtc_start = start + latency if start is not None else df.index.min()
    qtc_end = stop + latency if stop is not None else df.index.max()
    actual = trigger("Actual Start Trigger", "at most one")
    if actual is not None:
        if verbose:
            print("Found 'Actual Start Trigger' event, indicating cut-off of initial measurements. "
                  f"Will return actual start timestamp: {actual}")
        qtc_start = actual
    qtc_start, qtc_end = _as_utc(qtc_start), _as_utc(qtc_end)
    if verbose:
        print(f"EEG and EMG measurements last from {qtc_start} to {qtc_end}!\n")
    return qtc_start, qtc_end


def turn_trial_id_into_song_or_silence_id(log_df: pd.DataFrame, trial_id: int):
    """(song_id, silence_id) of a trial, one of them None (``data_integration.py:520-526``)."""
    first = log_df.loc[log_df["Trial ID"] == trial_id].iloc[0]
    song, silence = first["Song ID"], first["Silence ID"]
    return (int(song) if not np.isnan(song) else None), (int(silence) if not np.isnan(silence) else None)


def get_task_start_end(df: pd.DataFrame, song_id: int | None = None, song_title: str | None = None,
                       trial_id: int | None = None, silence_id: int | None = None,
                       assumed_latency_sec: float = 3.25, cut_off_sec_to_prevent_transients: float = 2.0,
                       verbose: bool = False):
    """(start, end) of one motor-task window: the rows of the trial (for music trials only those with a
    ``Task Frequency``), first / last timestamp + latency, end shortened by the transient cut-off; ``ValueError``
    for unknown, empty or excluded trials (``data_integration.py:604-714``)."""
    if song_id is None and song_title is None and silence_id is None and trial_id is None:
        raise ValueError("Either song_id, song_title, trial_id or silence_id must be specified")
    if trial_id is not None:
        song_id, silence_id = turn_trial_id_into_song_or_silence_id(df, trial_id)
    if song_id is not None or song_title is not None:
        if song_id is not None:
            subset = df.loc[df["Song ID"] == song_id]
        else:
            subset = df.loc[df["Song Title"] == song_title]
            ids = subset["Song ID"].dropna().unique().astype(int)
            if len(ids) > 1:
                raise ValueError(f"Song title appeared multiple times with Song IDs: {ids.tolist()}\n"
                                 "Choose one and call this method with song_id!")
        if verbose and subset["Song Skipped"].any():
            print(f"[INFO] Song {song_id} got skipped, no corresponding task was executed.")
        if verbose and subset["Trial Exclusion Bool"].any():
            print(f"[INFO] Song {song_id} marked for exclusion!")
        subset = subset.loc[~subset["Task Frequency"].isna()]
    else:
        subset = df.loc[df["Silence ID"] == silence_id]
        if verbose and subset["Trial Exclusion Bool"].any():
            print(f"[INFO] Silence trial {silence_id} marked for exclusion!")
    if len(subset) == 0:
        raise ValueError("Specific task not found!")
    if subset["Trial Exclusion Bool"].any():
        raise ValueError("Trial marked for exclusion!")
    if isinstance(subset.index, pd.DatetimeIndex):
        times = subset.index
    elif "Time" in subset.columns:
        times = pd.to_datetime(subset["Time"])
    else:
        raise ValueError('df must contain "Time" column or DatetimeIndex!')
    start, end = times.min(), times.max()
    if assumed_latency_sec > 0:
        start += pd.Timedelta(seconds=assumed_latency_sec)
        end += pd.Timedelta(seconds=assumed_latency_sec)
    if cut_off_sec_to_prevent_transients > 0:
        end = end - pd.Timedelta(seconds=cut_off_sec_to_prevent_transients)
    return start, end


def get_all_task_start_ends(enriched_log_df: pd.DataFrame, output_type: Literal["dict", "list"] = "dict",
                            assumed_latency_sec: float = 3.25, cut_off_sec_to_prevent_transients: float = 2.0):
    """Task windows of every valid trial, in order of first appearance; skipped / excluded / empty trials are left
    out (``data_integration.py:717-763``)."""
    spans_dict, spans_list = {}, []
    for trial in enriched_log_df["Trial ID"].unique():
        if pd.isna(trial):
            continue
        try:
            start, end = get_task_start_end(enriched_log_df, trial_id=trial, assumed_latency_sec=assumed_latency_sec,
                                            cut_off_sec_to_prevent_transients=cut_off_sec_to_prevent_transients)
        except ValueError:
            continue
        start, end = make_timezone_aware(start), make_timezone_aware(end)
        spans_dict[int(trial)] = (start, end)
        spans_list.append((start, end))
    return spans_dict if output_type == "dict" else spans_list
