"""The subject-condition sweeps (BASELINE config 5 shape at test size): pipelined results equal the item-by-item public
API, the fp64 oracle and the CBPA oracle; seeds are global so a rank's shard reproduces the single-process run."""
import numpy as np
import pytest
import torch
from scipy import signal

from oracle import cbpa as ocb
from oracle import coherence as oc
from oracle import surrogate as osur
from multimodal_biosignal_analysis_b200 import synthetic as syn

pytestmark = pytest.mark.gpu
FS, NPER = 512.0, 256


def _units(n_subj, conds, ne, nm, n_epochs=4):
    out = {}
    for s in range(n_subj):
        for c, cond in enumerate(conds):
            eeg, emg = syn.make_epochs(n_epochs, 1024, ne, nm, seed=900 + 10 * s + c)
            if cond == conds[0]:
                emg[:, :3] += 0.8 * eeg[:, :1]                     # condition A couples EEG 0 with EMG 0-2
            out[(f"S{s:02d}", cond)] = (eeg, emg)
    return out


@pytest.mark.parametrize("mode", ["phase", "shift"])
def test_surrogate_null_sweep_matches_item_by_item_and_oracle(cuda_device, mode):
    from multimodal_biosignal_analysis_b200 import data_surrogation as ds, signal_features as sf
    units = _units(3, ("A", "B"), 6, 8)
    recs = list(units.values())
    n_surr, seed = 96, 11
    got = []
    for res in ds.surrogate_null_sweep(recs, FS, nperseg=NPER, freq_band=(4, 60), n_surrogates=n_surr, mode=mode,
                                       seed=seed):
        got.append({k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in res.items()})
    assert [g["unit"] for g in got] == list(range(len(recs)))
    for u, (eeg, emg) in enumerate(recs):
        pc = sf.welch_magnitude_squared_coherence(eeg, emg, FS, nperseg=NPER, freq_band=(4, 60))
        if mode == "phase":
            ref = ds.phase_randomised_surrogate_null(pc, n_surr, seed=seed + u)
        else:
            ref = ds.circular_shift_surrogate_null(pc, n_surr, seed=seed + u)
        np.testing.assert_array_equal(got[u]["coherence"], pc.coherence)
        np.testing.assert_array_equal(got[u]["exceed"], ref["exceed"])
        np.testing.assert_array_equal(got[u]["max_stat"], ref["max_stat"])
        assert got[u]["threshold_fwe"] == ref["threshold_fwe"]
        np.testing.assert_allclose(got[u]["p_values"], ref["p_values"])
    # unit 2 against the fp64 definition
    eeg, emg = recs[2]
    starts = oc.welch_segments(eeg.shape[0], NPER)
    win = signal.get_window("hann", NPER)[None]
    f = np.fft.rfftfreq(NPER, 1 / FS)
    sel = np.flatnonzero((f >= 4) & (f <= 60))
    Xo = oc.segment_spectra(eeg, starts, win, 1, sel[0], sel[-1])[:, 0]
    Yo = oc.segment_spectra(emg, starts, win, 1, sel[0], sel[-1])[:, 0]
    assert np.max(np.abs(got[2]["coherence"] - oc.msc_from_spectra(Xo, Yo)[0])) < 1e-4
    Xw, _ = osur.whiten(Xo)
    Yw, _ = osur.whiten(Yo)
    shifts = np.random.default_rng(seed + 2).integers(1, len(starts), n_surr)
    cs = osur.surrogate_coherence(Xw, Yw, mode, np.arange(n_surr), shifts=shifts, seed=seed + 2)
    assert np.max(np.abs(got[2]["max_stat"] - cs.reshape(n_surr, -1).max(axis=1))) < 1e-4
    # a rank's shard with global unit indices reproduces its items
    part = list(ds.surrogate_null_sweep(recs[1::2], FS, nperseg=NPER, freq_band=(4, 60), n_surrogates=n_surr,
                                        mode=mode, seed=seed, unit_indices=range(1, len(recs), 2)))
    assert [p["unit"] for p in part] == [1, 3, 5]
    np.testing.assert_array_equal(part[-1]["exceed"], got[5]["exceed"])
    # numpy float64 (what np.load hands the reference's callers) gives the same result as float32
    r64 = next(ds.surrogate_null_sweep([(recs[0][0].astype(np.float64), recs[0][1].astype(np.float64))], FS,
                                       nperseg=NPER, freq_band=(4, 60), n_surrogates=n_surr, mode=mode, seed=seed))
    np.testing.assert_array_equal(r64["exceed"], got[0]["exceed"])


def test_full_sweep_contrast_and_cbpa_match_the_parts(cuda_device):
    """cmc_surrogate_cbpa_sweep == (coherence -> EMG-max -> A - B -> CBPA) assembled by hand, and the CBPA part is
    bit-exact against the MNE-algorithm oracle for the same sign table."""
    from multimodal_biosignal_analysis_b200 import cbpa as cb, signal_features as sf, sweep
    n_subj, ne, nm = 6, 8, 6
    units = _units(n_subj, ("A", "B"), ne, nm)
    adj_sp = cb.find_ch_adjacency_from_positions(syn.sensor_positions(64)[:ne])
    out = sweep.cmc_surrogate_cbpa_sweep(units, FS, nperseg=NPER, freq_band=(4, 60), n_surrogates=64, seed=5,
                                         spatial_adjacency=adj_sp, n_permutations=200, cbpa_seed=7,
                                         keep_pair_results=True)
    assert out["keys"] == list(units.keys()) and out["cmc"].shape[0] == 2 * n_subj
    for u, (key, (eeg, emg)) in enumerate(units.items()):
        coh = sf.welch_magnitude_squared_coherence(eeg, emg, FS, nperseg=NPER, freq_band=(4, 60)).coherence
        np.testing.assert_array_equal(out["cmc"][u], coh.max(axis=2))
        np.testing.assert_array_equal(out["pairs"][key]["coherence"], coh)
        assert out["n_significant_pairs"][u] == int((out["pairs"][key]["p_values"] < 0.05).sum())
    res = out["cbpa"][("A", "B")]
    X = np.stack([out["cmc"][2 * s].astype(np.float64) - out["cmc"][2 * s + 1].astype(np.float64)
                  for s in range(n_subj)])
    np.testing.assert_array_equal(res["X"], X)
    n_f = X.shape[1]
    adj = ocb.combine_adjacency(n_f, adj_sp)
    signs = cb.make_sign_table(200, n_subj, np.random.default_rng(7), 0)
    with np.errstate(all="ignore"):
        ref = ocb.permutation_cluster_1samp_test(X, signs, res["t_thresh"], 0, adj)
    np.testing.assert_array_equal(res["t_obs"], ref["t_obs"])
    np.testing.assert_array_equal(res["cluster_pv"], ref["cluster_pv"])
    np.testing.assert_array_equal(res["H0"], ref["H0_fixed"].astype(np.float64) / 2.0 ** 30)
    assert len(res["clusters"]) == len(ref["clusters"])
    for a, b in zip(res["clusters"], ref["clusters"]):
        np.testing.assert_array_equal(a, b)
    # the coupling planted in condition A shows up at EEG channel 0
    assert np.abs(res["t_obs"][:, 0]).max() > np.abs(res["t_obs"][:, 1:]).max()
    # masking by the surrogate p-values only ever removes coherence
    masked = sweep.cmc_surrogate_cbpa_sweep(units, FS, nperseg=NPER, freq_band=(4, 60), n_surrogates=64, seed=5,
                                            spatial_adjacency=adj_sp, n_permutations=50, mask_nonsignificant=True)
    assert np.all(masked["cmc"] <= out["cmc"]) and masked["cmc"].max() > 0
