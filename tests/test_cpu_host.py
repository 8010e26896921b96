"""CPU-only tests: the C-ABI library loads and exports every symbol declared in include/cmc.h, host-side
logic (sign tables, adjacency, sharding, file naming, the kept data_surrogation helpers) and the
world_size-2 sharding path over gloo.  No kernel is launched here."""
import os
import re
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from multimodal_biosignal_analysis_b200 import _lib, build
    build.build()
    hdr = open(os.path.join(ROOT, "include", "cmc.h")).read()
    declared = set(re.findall(r"CMC_API\s+[\w\s\*]+?\b(cmc_\w+)\s*\(", hdr))
    assert len(declared) >= 14
    lib = _lib.load()
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in cmc.h but not exported"
    assert declared == set(_lib.EXPORTED_SYMBOLS)
    assert lib.cmc_abi_version() == 1
    assert lib.cmc_csd_workspace_bytes(210, 100, 64, 64) > 0
    assert lib.cmc_csd_workspace_bytes(0, 100, 64, 64) < 0          # argument errors are negative codes


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "multimodal_biosignal_analysis_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert not re.search(r"^\s*(from|import)\s+oracle", src, re.M), fn


def test_no_cpu_fallback_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("needs a machine without a GPU")
    from multimodal_biosignal_analysis_b200 import signal_features as sf
    x = np.zeros((2048, 2), np.float32)
    with pytest.raises(RuntimeError):
        sf.multitaper_magnitude_squared_coherence(x, x, 256.0)
    with pytest.raises(ValueError):                                   # argument errors still come first
        sf.multitaper_magnitude_squared_coherence(x, x[:-1], 256.0)


def test_sign_table_random_and_exact():
    from multimodal_biosignal_analysis_b200.cbpa import make_sign_table
    s = make_sign_table(1000, 20, seed=42, tail=0)
    assert s.shape == (999, 20) and s.dtype == np.int8 and set(np.unique(s)) == {-1, 1}
    assert np.all(s[:, 0] == 1)                                       # two-tailed: subject 0 fixed
    np.testing.assert_array_equal(s, make_sign_table(1000, 20, seed=42, tail=0))
    e = make_sign_table(1000, 5, seed=1, tail=0)                      # 2**4 - 1 = 15 patterns: exact test
    assert e.shape == (15, 5) and len({tuple(r) for r in e}) == 15 and not np.any(np.all(e == 1, axis=1))
    e1 = make_sign_table(1000, 5, seed=1, tail=1)
    assert e1.shape == (31, 5)


def test_combine_adjacency_matches_oracle_and_wraparound():
    from multimodal_biosignal_analysis_b200 import cbpa, synthetic as syn
    from oracle import cbpa as ocb
    pos = syn.sensor_positions(64)
    sp = cbpa.find_ch_adjacency_from_positions(pos)
    assert (sp != ocb.delaunay_adjacency(pos)).nnz == 0
    a = cbpa.combine_adjacency(100, sp)
    assert a.shape == (6400, 6400)
    assert a.nnz == 2 * 99 * 64 + 100 * sp.nnz + 6400
    assert (a != ocb.combine_adjacency(100, sp)).nnz == 0
    assert (a != a.T).nnz == 0
    w = cbpa._add_phase_wraparound(a, 100, 64, np.arange(0, 360, 3.6))
    assert w[0 * 64 + 5, 99 * 64 + 5] and w[99 * 64 + 5, 5] and w.nnz == a.nnz + 128
    assert (w != ocb.add_phase_wraparound(a, 100, 64)).nnz == 0


def test_shard_range_partitions():
    from multimodal_biosignal_analysis_b200.dist import shard_range
    for n in (0, 1, 7, 1024, 10000):
        for ws in (1, 2, 3, 8):
            parts = [shard_range(n, r, ws) for r in range(ws)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(parts[i][1] == parts[i + 1][0] for i in range(ws - 1))
            sizes = [e - b for b, e in parts]
            assert max(sizes) - min(sizes) <= 1


def _gloo_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from multimodal_biosignal_analysis_b200 import dist as cd
    n = 11
    b, e = cd.shard_range(n)
    local = torch.arange(b, e, dtype=torch.int64) * 10
    full = cd.all_gather_ranges(local, n)
    cnt = torch.full((4,), rank + 1, dtype=torch.int32)
    cd.all_reduce_sum_(cnt)
    q.put((rank, full.tolist(), cnt.tolist()))
    dist.destroy_process_group()


def test_two_rank_sharding_over_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, full, cnt in out:
        assert full == [10 * i for i in range(11)]
        assert cnt == [3, 3, 3, 3]


def test_file_naming_and_spectrogram_roundtrip(tmp_path):
    from multimodal_biosignal_analysis_b200 import file_management as fm, signal_features as sf
    t = fm.file_title("Flexor CMC Spectrograms 11ch 1.00sec_step Channels_C3", ".npy")
    assert re.match(r"\d{4}-\d{2}-\d{2} \d{2}_\d{2}_\d{2} Flexor CMC Spectrograms", t) and t.endswith(".npy")
    spec = np.random.default_rng(0).random((5, 9, 11)).astype(np.float32)
    tc = np.arange(5) * 1.0 + 1.0
    fr = np.linspace(0, 4, 9)
    sf.save_spectrograms(spec, tc, fr, "Flexor CMC", tmp_path, "Channels_C3")
    names = sorted(os.listdir(tmp_path))
    assert any("Flexor CMC Spectrograms 11ch 1.00sec_step Channels_C3.npy" in n for n in names)
    assert any("Flexor CMC Timecenters 5windows Channels_C3.npy" in n for n in names)
    assert any("Flexor CMC Frequencies 9freqs Channels_C3.npy" in n for n in names)
    s2, t2, f2 = sf.fetch_stored_spectrograms(tmp_path, "Flexor CMC", ["Channels_C3"], expected_n_channels=11)
    np.testing.assert_array_equal(s2, spec)
    np.testing.assert_array_equal(t2, tc)
    np.testing.assert_array_equal(f2, fr)
    with pytest.raises(ValueError):
        sf.fetch_stored_spectrograms(tmp_path, "Flexor CMC", expected_n_channels=64)
    with pytest.raises(ValueError):
        fm.most_recent_file(tmp_path, ".npy", ["nothing matches"])


def test_insert_bad_channels_reference_properties():
    """Mirror of the reference's tests/test_data_surrogation.py:13-33."""
    from multimodal_biosignal_analysis_b200 import data_surrogation as ds
    x = np.random.randn(1000, 64)
    assert np.array_equal(x, ds.insert_bad_channels(x, axis=0, scale_range=(1.0, 1.0))[0])
    out, amended = ds.insert_bad_channels(x, axis=0)
    keep = [i for i in range(64) if i + 1 not in amended]
    assert np.array_equal(x[:, keep], out[:, keep]) and not np.array_equal(x, out)
    assert len(amended) == 5
    with pytest.raises(AttributeError):
        ds.insert_bad_channels(x)                                    # 2-D input without axis
    noisy = ds.add_noise_to_channels(x, 0.0, [1, 3], axis=0, noise_type="pink", random_seed=3)
    assert np.array_equal(noisy[:, 0], x[:, 0]) and not np.array_equal(noisy[:, 1], x[:, 1])
    snr = 10 * np.log10(np.mean(x[:, 1] ** 2) / np.mean((noisy[:, 1] - x[:, 1]) ** 2))
    assert abs(snr) < 1e-9
    with pytest.raises(ValueError):
        ds.add_noise_to_channels(x, 0.0, [64], axis=0)


def test_scalar_statistics_match_reference_golden():
    from conftest import golden
    from multimodal_biosignal_analysis_b200 import signal_features as sf
    g = golden("scalars.npz")
    assert sf.compute_cmc_independence_threshold(5, 0.05) == float(g["IT_K5_a05"])
    np.testing.assert_allclose(sf.fisher_atanh_transform(g["fisher_in"]), g["fisher_out"], rtol=1e-15)
    np.testing.assert_allclose(sf.inverse_fisher_atanh(g["inv_in"]), g["inv_out"], rtol=1e-15)
    m, it = sf.apply_threshold_filtering(np.array([0.5, 0.9]), K=5, alpha=0.05)
    assert it == float(g["IT_K5_a05"]) and m.tolist() == [False, True]
    m, it2 = sf.apply_threshold_filtering(np.array([0.5, 0.9]), K=5, alpha=0.05, n_comparisons=4096,
                                          apply_bonferroni=True)
    assert it2 > it
    g2 = golden("max_over_emg.npz")
    a, b, d = sf.max_cmc_spectrograms_over_channels(g2["c"], g2["lo"], g2["hi"], verbose=False)
    np.testing.assert_array_equal(a, g2["a"])
    np.testing.assert_array_equal(b, g2["b"])
    np.testing.assert_array_equal(d, g2["d"])


def test_cbpa_config_defaults_match_reference():
    from multimodal_biosignal_analysis_b200.cbpa import CBPAConfig, CMC_CHANNEL_FILE_SUFFIX
    c = CBPAConfig()
    assert (c.alpha_cluster_forming, c.n_permutations, c.tail, c.use_spatio_temporal, c.n_jobs, c.seed) == \
        (0.05, 1000, 0, True, -1, 42)
    assert (c.use_phase_normalization, c.n_phase_bins, c.cmc_time_window_sec, c.psd_time_window_sec) == \
        (False, 36, 2.0, 0.25)
    assert CMC_CHANNEL_FILE_SUFFIX == "Channels_C5_C3_C1_FC5_FC3_FC1_F3_CP5_CP3_CP1_P3"


def test_bench_reference_arm_helpers():
    """the CPU arm of bench.py agrees with the oracle it is built from (tiny case, single process)."""
    sys.path.insert(0, ROOT)
    import bench
    from oracle import coherence as oc
    from multimodal_biosignal_analysis_b200 import synthetic as syn
    eeg, emg = syn.make_epochs(1, 8192, 4, 3, seed=1)
    starts = syn.epoch_segment_starts(1, 8192, 2048, 1024)
    c = bench.cpu_pooled_coherence(eeg.astype(np.float64), emg.astype(np.float64), starts, None, 1)
    ref = oc.welch_msc(eeg[:8192], emg[:8192], 2048, bin_lo=1, bin_hi=100)
    np.testing.assert_allclose(c, ref, atol=1e-12)


def test_spectrogram_aggregation_matches_reference_golden():
    from conftest import golden
    from multimodal_biosignal_analysis_b200 import signal_features as sf
    g = golden("aggregation.npz")
    spec, fr = g["spec"], g["freqs"]
    bands = {'alpha': (8, 12), 'beta': (13, 30)}
    for beh in ('mean', 'max'):
        d = sf.aggregate_spectrogram_over_frequency_band(spec, fr, behaviour=beh, frequency_bands=bands,
                                                         lower_array=spec * 0.8, upper_array=spec * 1.2)
        for b in bands:
            for k, nm in enumerate(('v', 'lo', 'hi')):
                np.testing.assert_array_equal(d[b][k], g[f"band_{beh}_{b}_{nm}"])
    d = sf.aggregate_spectrogram_over_frequency_band(spec, fr, behaviour='mean', frequency_bands=bands,
                                                     log_transform=True, pre_aggregate_axis=(2, 'max'))
    np.testing.assert_array_equal(d['alpha'], g["band_pre_alpha"])
    np.testing.assert_array_equal(
        sf.aggregate_psd_spectrogram(spec, fr, normalize_mvc=True, freq_slice='slow',
                                     aggregation_ops=[('mean', 1), ('max', 1)]), g["psd_agg_emg"])
    np.testing.assert_array_equal(
        sf.aggregate_psd_spectrogram(spec, fr, channel_indices=[0, 1, 4], freq_slice=(8, 12),
                                     aggregation_ops=[('mean', 2), ('mean', 1)]), g["psd_agg_eeg"])
    # the intended masking is available explicitly and differs from the reference's np.take behaviour
    strict = sf.aggregate_spectrogram_over_frequency_band(spec, fr, frequency_bands=bands, strict_band_mask=True)
    sel = (fr >= 8) & (fr < 12)
    np.testing.assert_allclose(strict['alpha'], spec[:, sel].mean(axis=1))
    with pytest.raises(ValueError):
        sf.aggregate_spectrogram_over_frequency_band(spec, fr[:-1])
    with pytest.raises(ValueError):
        sf.aggregate_psd_spectrogram(spec, None, freq_slice='alpha')
