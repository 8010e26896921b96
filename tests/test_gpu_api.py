"""GPU tests of the reference-facing Python surface (numpy in -> numpy out), mirroring how the
reference's own callers and tests use it."""
import numpy as np
import pandas as pd
import pytest
import torch
from scipy.stats import t as t_dist

from conftest import golden
from oracle import cbpa as ocb
from oracle import coherence as oc
from multimodal_biosignal_analysis_b200 import synthetic as syn

pytestmark = pytest.mark.gpu


def test_multitaper_msc_dict_matches_reference_golden(cuda_device):
    from multimodal_biosignal_analysis_b200 import signal_features as sf
    g = golden("msc_jk.npz")
    r = sf.multitaper_magnitude_squared_coherence(
        g["eeg"], g["emg"], float(g["fs"]), nw=3, window_length_sec=1.0, overlap_frac=0.5, use_jackknife=True,
        jackknife_alpha=0.05, apply_independence_threshold=True, apply_bonferroni_correction=True,
        significance_level=0.2, window_mask=g["window_mask"])
    assert isinstance(r["coherence_raw"], np.ndarray) and r["coherence_raw"].dtype == np.float32
    assert r["coherence_significant"].dtype == bool
    for k, gk in (("coherence_raw", "coherence_raw"), ("coherence_ci_lower", "ci_lower"),
                  ("coherence_ci_upper", "ci_upper")):
        assert r[k].shape == g[gk].shape and np.max(np.abs(r[k] - g[gk])) < 1e-4
    np.testing.assert_array_equal(r["time_centers"], g["time_centers"])
    np.testing.assert_array_equal(r["freqs"], g["freqs"])
    assert r["metadata"]["K_tapers"] == 5 and r["metadata"]["n_active_windows"] == int(g["window_mask"].sum())
    assert abs(r["metadata"]["IT_bonferroni"] - float(g["IT_bonferroni"])) < 1e-12
    assert r["metadata"]["n_significant"] == int(r["coherence_significant"].sum())


def test_multitaper_msc_axis_chunking_and_errors(cuda_device, monkeypatch):
    from multimodal_biosignal_analysis_b200 import signal_features as sf
    g = golden("msc_axis_overlap.npz")
    kw = dict(nw=3, window_length_sec=0.5, overlap_frac=0.75, use_jackknife=False,
              apply_independence_threshold=False)
    r = sf.multitaper_magnitude_squared_coherence(g["eeg"].T.copy(), g["emg"], float(g["fs"]), eeg_axis=1, **kw)
    assert np.max(np.abs(r["coherence_raw"] - g["coherence_raw"])) < 1e-4
    monkeypatch.setattr(sf, "DEVICE_CHUNK_BYTES", 20000)          # force several window chunks
    r2 = sf.multitaper_magnitude_squared_coherence(g["eeg"], g["emg"], float(g["fs"]), **kw)
    np.testing.assert_array_equal(r2["coherence_raw"], r["coherence_raw"])
    # CUDA tensors in -> CUDA tensors out
    r3 = sf.multitaper_magnitude_squared_coherence(torch.as_tensor(g["eeg"]).cuda(), torch.as_tensor(g["emg"]).cuda(),
                                                   float(g["fs"]), **kw)
    assert r3["coherence_raw"].is_cuda
    np.testing.assert_array_equal(r3["coherence_raw"].cpu().numpy(), r2["coherence_raw"])
    with pytest.raises(ValueError, match="same number of samples"):
        sf.multitaper_magnitude_squared_coherence(g["eeg"][:-1], g["emg"], float(g["fs"]))
    with pytest.raises(ValueError, match="window_mask must have shape"):
        sf.multitaper_magnitude_squared_coherence(g["eeg"], g["emg"], float(g["fs"]), window_mask=np.ones(3, bool))
    with pytest.raises(ValueError, match="must be 2D"):
        sf.multitaper_magnitude_squared_coherence(g["eeg"][:, 0], g["emg"], float(g["fs"]))
    # band-limited + fused EMG reduction agree with the full result
    full = sf.multitaper_magnitude_squared_coherence(g["eeg"], g["emg"], float(g["fs"]), nw=3, window_length_sec=0.5,
                                                     overlap_frac=0.75, use_jackknife=True,
                                                     apply_independence_threshold=False)
    red = sf.multitaper_magnitude_squared_coherence(g["eeg"], g["emg"], float(g["fs"]), nw=3, window_length_sec=0.5,
                                                    overlap_frac=0.75, use_jackknife=True,
                                                    apply_independence_threshold=False, freq_band=(8, 40),
                                                    reduce_emg=True)
    sel = (full["freqs"] >= 8) & (full["freqs"] <= 40)
    a, b, c = sf.max_cmc_spectrograms_over_channels(full["coherence_raw"][:, sel], full["coherence_ci_lower"][:, sel],
                                                    full["coherence_ci_upper"][:, sel], verbose=False)
    np.testing.assert_array_equal(red["coherence_raw"], a)
    np.testing.assert_array_equal(red["coherence_ci_lower"], b)
    np.testing.assert_array_equal(red["coherence_ci_upper"], c)
    np.testing.assert_array_equal(red["freqs"], full["freqs"][sel])


def test_jackknife_helper_matches_reference_golden(cuda_device):
    from multimodal_biosignal_analysis_b200 import signal_features as sf
    g = golden("jackknife_window.npz")
    m, lo, hi = sf.jackknife_coherence_and_ci(list(g["tapers"]), g["eeg"], g["emg"], float(g["fs"]), 128,
                                              jackknife_alpha=0.1)
    assert np.max(np.abs(m - g["mean"])) < 1e-4
    assert np.max(np.abs(lo - g["lower"])) < 1e-4
    assert np.max(np.abs(hi - g["upper"])) < 1e-4


def test_taskwise_cmc_buffer_independence_and_real_path(cuda_device, monkeypatch):
    """Port of the reference's tests/test_signal_features.py:252-329 (mask geometry), then the real
    fused path against the oracle."""
    from multimodal_biosignal_analysis_b200 import signal_features as features
    t0 = pd.Timestamp("2026-01-01 00:00:00", tz="UTC")
    trial = (t0 + pd.Timedelta(seconds=10), t0 + pd.Timedelta(seconds=20))
    monkeypatch.setattr(features.data_integration, "get_all_task_start_ends", lambda *a, **k: [trial])
    monkeypatch.setattr(features.data_integration, "get_qtc_measurement_start_end",
                        lambda *a, **k: (t0, t0 + pd.Timedelta(seconds=30)))

    def fake(subset_eeg, subset_emg, sampling_freq, window_length_sec, overlap_frac, **_kw):
        ws = int(window_length_sec * sampling_freq)
        hop = int(ws * (1 - overlap_frac))
        n = (len(subset_eeg) - ws) // hop + 1
        c = np.zeros((n, 1, 1, 1), np.float32)
        for w in range(n):
            c[w, 0, 0, 0] = float(np.mean(subset_eeg[w * hop:w * hop + ws, 0]))
        return {"coherence_raw": c, "time_centers": (np.arange(n) * hop + ws / 2) / sampling_freq,
                "freqs": np.array([0.0]), "metadata": {}}

    eeg = np.arange(31, dtype=np.float32).reshape(-1, 1)
    emg = np.zeros((31, 1), np.float32)
    with monkeypatch.context() as mp:
        mp.setattr(features, "multitaper_magnitude_squared_coherence", fake)
        mp.setattr(features, "max_cmc_spectrograms_over_channels", lambda cmc, *a, **k: cmc[..., 0])
        results = {}
        for buf in (0.0, 1.0, 3.0, 5.0):
            vals, tc, _ = features.compute_task_wise_aggregated_cmc(
                eeg_array=eeg, emg_array=emg, sampling_freq=1, muscle_group="test",
                log_frame=pd.DataFrame({"dummy": [1]}), window_size_sec=2.0, window_overlap_ratio=0.5,
                use_jackknife=False, pre_trial_computation_buffer_sec=buf, post_trial_computation_buffer_sec=buf)
            results[buf] = (vals, tc)
        ref_vals, ref_t = results[0.0]
        core = (ref_t >= 12.0) & (ref_t < 18.0)
        for buf in (1.0, 3.0, 5.0):
            v, t = results[buf]
            assert np.nanmax(np.abs(ref_vals[core] - v[(t >= 12.0) & (t < 18.0)])) == pytest.approx(0.0, abs=1e-7)

    # real path: 30 s at 256 Hz, trial 10-20 s, fused EMG-argmax with jackknife CI
    fs = 256
    e, m = syn.make_recording(30 * fs, 6, 5, seed=4, fs=float(fs))
    vals, lo, hi, tc, fr = features.compute_task_wise_aggregated_cmc(
        e, m, fs, "flexor", log_frame=pd.DataFrame({"dummy": [1]}), window_size_sec=1.0,
        window_overlap_ratio=0.5, use_jackknife=True, pre_trial_computation_buffer_sec=1.0,
        post_trial_computation_buffer_sec=1.0)
    assert vals.shape == (59, 129, 6) and lo.shape == vals.shape and hi.shape == vals.shape
    mask = (tc >= 9.0) & (tc <= 21.0)
    assert np.all(vals[~mask] == 0) and np.all(vals[mask].max(axis=(1, 2)) > 0)
    o = oc.multitaper_msc(e, m, float(fs), window_length_sec=1.0, use_jackknife=True,
                          apply_independence_threshold=False, window_mask=mask)
    a, b, c = oc.max_over_emg(o["coherence_raw"], o["coherence_ci_lower"], o["coherence_ci_upper"])
    assert np.max(np.abs(vals - a)) < 1e-4
    assert np.all(lo <= vals) and np.all(vals <= hi)            # the reference's CI sanity asserts (:986-990)
    # gathered CI may come from a different EMG channel only where two channels tie within tolerance
    assert np.mean(np.abs(lo - b) < 1e-4) > 0.999 and np.mean(np.abs(hi - c) < 1e-4) > 0.999


def test_welch_api_and_surrogate_api(cuda_device):
    from multimodal_biosignal_analysis_b200 import signal_features as sf, data_surrogation as ds
    eeg, emg = syn.make_epochs(4, 4096, 8, 12, seed=9)
    starts = syn.epoch_segment_starts(4, 4096, 1024, 512)
    pc = sf.welch_magnitude_squared_coherence(eeg, emg, 2048.0, nperseg=1024, freq_band=(2, 60), segment_starts=starts)
    ref = oc.welch_msc(eeg, emg, 1024)            # regular grid differs from the epoch grid: recompute below
    win = __import__("scipy.signal", fromlist=["x"]).get_window("hann", 1024)[None]
    Xo = oc.segment_spectra(eeg, starts, win, 1, 1, 30)[:, 0]
    Yo = oc.segment_spectra(emg, starts, win, 1, 1, 30)[:, 0]
    coh = oc.msc_from_spectra(Xo, Yo)[0]
    assert pc.coherence.shape == (30, 8, 12) and np.max(np.abs(pc.coherence - coh)) < 1e-4
    np.testing.assert_allclose(pc.freqs, np.arange(1, 31) * 2.0)
    del ref
    with pytest.warns(RuntimeWarning, match="distinct shifts"):      # 28 segments: only 27 distinct surrogates exist
        null = ds.circular_shift_surrogate_null(pc, 200, seed=5)
    assert null["n_distinct_shifts"] <= 27 and abs(null["p_resolution"] - 1 / 28) < 1e-12
    assert null["exceed"].shape == coh.shape and null["max_stat"].shape == (200,)
    assert np.all(null["p_values"] > 0) and np.all(null["p_values"] <= 1)
    assert 0 < null["threshold_fwe"] <= 1
    nullp = ds.phase_randomised_surrogate_null(pc, 256, seed=5)
    assert nullp["exceed"].max() <= 256 and nullp["max_stat"].shape == (256,)
    # strongly coupled pairs are significant under both nulls, uncoupled ones are not systematically
    strong = coh.max(axis=0) > 0.5
    if strong.any():
        assert np.all(nullp["p_values"].min(axis=0)[strong] < 0.02)


def test_cbpa_api_matches_oracle_and_result_dict(cuda_device, tmp_path):
    from multimodal_biosignal_analysis_b200 import cbpa as cb
    X = syn.make_cbpa_contrast(13, 36, 11, seed=6)
    pos = syn.sensor_positions(64)[:11]
    cfg = cb.CBPAConfig(modality="CMC", freq_band="beta", n_permutations=64, tail=0, use_phase_normalization=True,
                        output_dir=tmp_path, hypothesis_label="unit", save_plots=False)
    res = cb.run_cbpa(cfg, contrast=(X, [f"ch{i}" for i in range(11)], np.arange(36) * 10.0),
                      spatial_adjacency=pos)
    for key in ("t_obs", "t_thresh", "clusters", "cluster_pv", "H0", "good_cluster_inds", "ch_names", "time_grid",
                "cfg", "n_valid_subjects"):
        assert key in res
    assert res["t_obs"].shape == (36, 11) and res["H0"].shape == (64,) and res["n_valid_subjects"] == 13
    assert all(c.dtype == bool and c.shape == (36, 11) for c in res["clusters"])
    adj = ocb.add_phase_wraparound(ocb.combine_adjacency(36, ocb.delaunay_adjacency(pos)), 36, 11)
    signs = cb.make_sign_table(64, 13, np.random.default_rng(42), 0)
    ref = ocb.permutation_cluster_1samp_test(X, signs, res["t_thresh"], 0, adj)
    np.testing.assert_array_equal(res["t_obs"], ref["t_obs"])
    np.testing.assert_array_equal(res["H0"], ref["H0"])
    np.testing.assert_array_equal(res["cluster_pv"], ref["cluster_pv"])
    assert len(res["clusters"]) == len(ref["clusters"])
    for a, b in zip(res["clusters"], ref["clusters"]):
        np.testing.assert_array_equal(a, b)
    files = sorted(p.name for p in tmp_path.iterdir())
    assert any(f.endswith("unit.npz") for f in files) and any(f.endswith("_t_obs.csv") for f in files)
    assert any(f.endswith("_cluster_summary.csv") for f in files)
    with pytest.raises(ValueError, match="incompatible tail"):
        cb.permutation_cluster_1samp_test(X, threshold=-2.0, tail=1, adjacency=adj)


def test_multitaper_psd_matches_reference_golden(cuda_device):
    from multimodal_biosignal_analysis_b200 import signal_features as sf
    g = golden("psd.npz")
    s_log, tc, fr = sf.multitaper_psd(g["x"], float(g["fs"]), nw=3, window_length_sec=0.5, overlap_frac=0.5, axis=0,
                                      apply_log_scale=True)
    s_lin, _, _ = sf.multitaper_psd(g["x"], float(g["fs"]), nw=3, window_length_sec=0.5, overlap_frac=0.5, axis=0,
                                    apply_log_scale=False)
    assert s_log.shape == g["s_log"].shape                     # the PSD grid has one window fewer than the MSC grid
    np.testing.assert_array_equal(tc, g["time_centers"])
    np.testing.assert_array_equal(fr, g["freqs"])
    # DC is exactly 0 after the post-taper mean removal; elsewhere float32 relative accuracy
    assert np.all(s_lin[:, 0] == 0)
    np.testing.assert_allclose(s_lin[:, 1:], g["s_lin"][:, 1:], rtol=2e-4, atol=1e-9)
    assert np.max(np.abs(s_log[:, 1:] - g["s_log"][:, 1:])) < 1e-4
    with pytest.raises(AttributeError):
        sf.multitaper_psd(g["x"], float(g["fs"]))             # 2-D input without axis


def test_spectral_snr_and_welch_psd_vs_scipy(cuda_device):
    """Includes the reference's tests/test_signal_features.py:14-24 (scale invariance) on its own shapes:
    1000 samples at 500 Hz -> nperseg clipped to 1000, a non power-of-two length (direct DFT kernel)."""
    from scipy import signal as ss
    from multimodal_biosignal_analysis_b200 import signal_features as sf
    rng = np.random.default_rng(0)
    x = rng.standard_normal((1000, 64))
    a = sf.compute_spectral_snr(x, 500)
    b = sf.compute_spectral_snr(x * .5, 500)
    assert a == pytest.approx(b, abs=1e-4)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        fr, ref = ss.welch(x, axis=0, fs=500, nperseg=2000)
    snr, freqs, psd = sf.compute_spectral_snr(x, 500, return_psd=True)
    np.testing.assert_allclose(freqs, fr)
    np.testing.assert_allclose(psd, ref, rtol=3e-4, atol=1e-9)
    tb = (fr < 21.5 + 4.25) & (fr > 21.5 - 4.25)
    nb = (fr >= 13.0) & (fr <= 30.0)
    assert snr == pytest.approx(10 * np.log10(np.mean(ref[tb]) / np.mean(ref[nb])), abs=1e-3)
    # power-of-two path: 2048 Hz, 4-s segments of 8192 samples
    y = rng.standard_normal((8192 * 3, 5)) + 0.3
    f2, p2 = sf.welch_psd(y, 2048, 8192)
    fr2, ref2 = ss.welch(y, axis=0, fs=2048, nperseg=8192)
    np.testing.assert_allclose(p2, ref2, rtol=3e-4, atol=1e-12)
    r = sf.resample_data(x, 500, 1000, axis=0)
    assert r.shape == (2000, 64)


def test_local_neighbor_coherence_vs_scipy_loop(cuda_device):
    from scipy import signal as ss
    from multimodal_biosignal_analysis_b200 import signal_features as sf
    rng = np.random.default_rng(2)
    common = rng.standard_normal(6000)
    data = rng.standard_normal((6000, 6)) + common[:, None] * np.array([0.0, 0.5, 1.0, 1.0, 0.2, 0.0])
    mapping = [[1, 2], [0, 2], [1, 3], [2, 4], [3, 5], [4]]
    ref = []
    for ch, nbs in enumerate(mapping):
        ref.append(np.nanmean([np.nanmean(ss.coherence(data[:, ch], data[:, nb], fs=512.0)[1]) for nb in nbs]))
    got = sf.local_neighbor_coherence(data, mapping, 512.0)
    assert got == pytest.approx(float(np.nanmean(ref)), abs=0.01)       # DC bin: 0/0 in both, 1/129 of the mean


def test_welch_coherence_sweep_matches_item_by_item(cuda_device):
    """The three-stream sweep (upload / K1 + K2 / download overlapped across recordings) returns, item by item,
    exactly what welch_magnitude_squared_coherence returns - for pinned tensors, plain numpy arrays and a
    single-item sweep, with odd channel counts."""
    import torch
    from multimodal_biosignal_analysis_b200 import signal_features as sf, synthetic as syn
    fs, nper = 512.0, 256
    recs = []
    for k in range(5):
        eeg, emg = syn.make_epochs(2, 2048, 7, 10, seed=60 + k)
        recs.append((eeg, emg))
    ref = [sf.welch_magnitude_squared_coherence(e, m, fs, nperseg=nper, freq_band=(2, 60)).coherence for e, m in recs]
    pinned = [(torch.from_numpy(e).pin_memory(), torch.from_numpy(m).pin_memory()) for e, m in recs]
    for items in (pinned, recs, recs[:1], (r for r in pinned[:2])):
        got = []
        for coh, freqs in sf.welch_coherence_sweep(items, fs, nperseg=nper, freq_band=(2, 60)):
            got.append(coh.copy())
            assert freqs[0] >= 2 and freqs[-1] <= 60 and coh.shape == (len(freqs), 7, 10)
        assert len(got) >= 1
        for g, r in zip(got, ref):
            np.testing.assert_array_equal(g, r)
    assert list(sf.welch_coherence_sweep([], fs)) == []
    # documented lifetime: item i stays valid while item i + 1 is fetched (no copy taken here)
    held = None
    for k, (coh, _) in enumerate(sf.welch_coherence_sweep(pinned, fs, nperseg=nper, freq_band=(2, 60))):
        if held is not None:
            torch.cuda.synchronize()                              # everything submitted so far has landed
            np.testing.assert_array_equal(held, ref[k - 1])
        held = coh


def test_sharded_nulls_and_cbpa_identical_on_two_gpus(cuda_device):
    """torchrun with 2 ranks (NCCL): frequency-sharded and surrogate-sharded nulls and the index-sharded CBPA
    return exactly the single-GPU results (scripts/check_multi_gpu.py).  Needs two GPUs."""
    import os
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", str(29600 + os.getpid() % 300), os.path.join(root, "scripts", "check_multi_gpu.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "identical to one GPU: True" in r.stdout


def test_welch_coherence_sweep_can_be_abandoned(cuda_device):
    """Stopping the sweep early (generator closed with an item still in flight) drains its streams before the
    buffers go away; a following call is unaffected."""
    import torch
    from multimodal_biosignal_analysis_b200 import signal_features as sf, synthetic as syn
    recs = [syn.make_epochs(2, 2048, 6, 8, seed=80 + k) for k in range(4)]
    ref = sf.welch_magnitude_squared_coherence(*recs[0], 512.0, nperseg=256).coherence
    gen = sf.welch_coherence_sweep(recs, 512.0, nperseg=256)
    first, _ = next(gen)
    first = first.copy()
    gen.close()
    torch.cuda.synchronize()
    np.testing.assert_array_equal(first, ref)
    np.testing.assert_array_equal(sf.welch_magnitude_squared_coherence(*recs[0], 512.0, nperseg=256).coherence, ref)


def test_run_batch_unchanged_call_on_stored_files(cuda_device, tmp_path):
    """The call of the reference's statistics workflow (statistics_RQ_A_post_hoc_testing_workflow.py:136-172, 465):
    ``run_batch(CONTRASTS)`` with nothing but CBPAConfig objects.  The contrast comes from stored spectrogram files
    (tests/cbpa_fixture.py; X is pinned to the reference by tests/test_cpu_host.py), the channel adjacency from the
    cap layout, the permutations run on the GPU; result dict, .npz and CSVs follow cbpa.py:1051-1056, 1076-1185 and
    the statistics equal the MNE-algorithm oracle for the same sign table."""
    import pandas as pd
    import cbpa_fixture as fx
    from oracle import cbpa as ocb
    from multimodal_biosignal_analysis_b200 import cbpa as cb
    root = fx.build_study(tmp_path / "study")
    cfgs = fx.configs(cb, root, tmp_path / "out")
    contrasts = [cfgs["cmc_clock"], cfgs["cmc_phase"]]
    all_results, combined = cb.run_batch(contrasts)
    assert len(all_results) == 2
    for cfg, res in zip(contrasts, all_results):
        assert set(res) == {"t_obs", "t_thresh", "clusters", "cluster_pv", "H0", "good_cluster_inds", "ch_names",
                            "time_grid", "cfg", "n_valid_subjects"}
        X, ch, grid = cb.build_contrast_array(cfg)
        n_subj, n_times, n_ch = X.shape
        assert res["n_valid_subjects"] == n_subj == 4 and res["t_obs"].shape == (n_times, n_ch)
        # 4 subjects, two-tailed: all 2^3 - 1 = 7 sign patterns fit into 64 permutations -> exact test, like MNE
        assert list(res["ch_names"]) == fx.CMC_SUBSET and res["H0"].shape == (8,)
        adj = cb.combine_adjacency(n_times, cb.default_spatial_adjacency(fx.CMC_SUBSET))
        if cfg.use_phase_normalization:
            adj = cb._add_phase_wraparound(adj, n_times, n_ch, np.asarray(grid))
        signs = cb.make_sign_table(cfg.n_permutations, n_subj, np.random.default_rng(cfg.seed), cfg.tail)
        with np.errstate(all="ignore"):
            ref = ocb.permutation_cluster_1samp_test(X, signs, float(res["t_thresh"]), cfg.tail, adj)
        np.testing.assert_array_equal(res["t_obs"], ref["t_obs"])
        np.testing.assert_array_equal(res["cluster_pv"], ref["cluster_pv"])
        np.testing.assert_array_equal(res["H0"], ref["H0"])
        assert len(res["clusters"]) == len(ref["clusters"])
        for a, b in zip(res["clusters"], ref["clusters"]):
            np.testing.assert_array_equal(a, b)
    out = tmp_path / "out"
    names = sorted(p.name for p in out.iterdir())
    assert sum(n.endswith("cmc clock.npz") for n in names) == 1 and sum(n.endswith("cmc phase_t_obs.csv") for n in names) == 1
    comb = [n for n in names if "CBPA Combined Cluster Summary" in n]
    assert len(comb) == 1
    df = pd.read_csv(out / comb[0])
    assert len(df) == sum(len(r["clusters"]) for r in all_results) == len(combined)
    assert {"hypothesis", "cluster_index", "p_value", "significant", "peak_t", "n_channels", "channels"} <= set(df.columns)
    # run_cbpa(cfg) alone writes its own per-run cluster summary
    cb.run_cbpa(cfgs["psd_alpha"])
    assert any(n.name.endswith("psd alpha_cluster_summary.csv") for n in out.iterdir())
