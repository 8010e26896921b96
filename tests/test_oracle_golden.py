"""Pin the numpy oracle against outputs of the reference's own functions
(tests/golden/*.npz, produced by scripts/make_golden.py) and against scipy."""
import numpy as np

from conftest import golden
from oracle import coherence as oc


def test_msc_nojk_matches_reference():
    g = golden("msc_nojk.npz")
    r = oc.multitaper_msc(g["eeg"], g["emg"], float(g["fs"]), use_jackknife=False,
                          apply_independence_threshold=True, significance_level=0.05)
    assert r["coherence_raw"].shape == g["coherence_raw"].shape
    assert np.max(np.abs(r["coherence_raw"] - g["coherence_raw"])) < 2e-7
    np.testing.assert_array_equal(r["time_centers"], g["time_centers"])
    np.testing.assert_array_equal(r["freqs"], g["freqs"])
    assert r["metadata"]["K_tapers"] == int(g["K"])
    assert abs(r["metadata"]["IT_unadjusted"] - float(g["IT"])) < 1e-15
    # the mask may differ only where the coherence sits within float32 eps of IT
    diff = r["coherence_significant"] != g["coherence_significant"]
    assert np.all(np.abs(g["coherence_raw"][diff] - float(g["IT"])) < 1e-6)


def test_msc_jackknife_mask_bonferroni_matches_reference():
    g = golden("msc_jk.npz")
    r = oc.multitaper_msc(g["eeg"], g["emg"], float(g["fs"]), use_jackknife=True,
                          jackknife_alpha=0.05, apply_independence_threshold=True,
                          apply_bonferroni_correction=True, significance_level=0.2,
                          window_mask=g["window_mask"])
    # the reference accumulates the jackknife in float32/complex64 (:503-510)
    for k_or, k_g in (("coherence_raw", "coherence_raw"), ("coherence_ci_lower", "ci_lower"),
                      ("coherence_ci_upper", "ci_upper")):
        assert np.max(np.abs(r[k_or] - g[k_g])) < 2e-5, k_or
    skipped = ~g["window_mask"]
    assert np.all(r["coherence_raw"][skipped] == 0)
    assert np.all(g["time_centers"] == r["time_centers"])
    diff = r["coherence_significant"] != g["coherence_significant"]
    assert np.all(np.abs(g["coherence_raw"][diff] - float(g["IT_bonferroni"])) < 1e-5)


def test_jackknife_window_matches_reference():
    g = golden("jackknife_window.npz")
    X = oc.segment_spectra(g["eeg"], np.array([0]), g["tapers"])[0]
    Y = oc.segment_spectra(g["emg"], np.array([0]), g["tapers"])[0]
    m, lo, hi = oc.jackknife_from_spectra(X, Y, alpha=0.1)
    assert np.max(np.abs(m - g["mean"])) < 2e-5
    assert np.max(np.abs(lo - g["lower"])) < 2e-5
    assert np.max(np.abs(hi - g["upper"])) < 2e-5


def test_msc_other_overlap_matches_reference():
    g = golden("msc_axis_overlap.npz")
    r = oc.multitaper_msc(g["eeg"], g["emg"], float(g["fs"]), window_length_sec=0.5,
                          overlap_frac=0.75, use_jackknife=False,
                          apply_independence_threshold=False)
    assert r["coherence_raw"].shape == g["coherence_raw"].shape
    assert np.max(np.abs(r["coherence_raw"] - g["coherence_raw"])) < 2e-7
    np.testing.assert_array_equal(r["time_centers"], g["time_centers"])


def test_max_over_emg_matches_reference():
    g = golden("max_over_emg.npz")
    a, b, d = oc.max_over_emg(g["c"], g["lo"], g["hi"])
    np.testing.assert_array_equal(a, g["a"])
    np.testing.assert_array_equal(b, g["b"])
    np.testing.assert_array_equal(d, g["d"])


def test_psd_matches_reference():
    g = golden("psd.npz")
    s_log, tc, fr = oc.multitaper_psd(g["x"], float(g["fs"]), window_length_sec=0.5,
                                      apply_log_scale=True)
    s_lin, _, _ = oc.multitaper_psd(g["x"], float(g["fs"]), window_length_sec=0.5,
                                    apply_log_scale=False)
    assert s_log.shape == g["s_log"].shape
    np.testing.assert_allclose(s_lin, g["s_lin"], rtol=1e-10, atol=1e-18)
    np.testing.assert_allclose(s_log, g["s_log"], rtol=0, atol=1e-10)
    np.testing.assert_array_equal(tc, g["time_centers"])
    np.testing.assert_array_equal(fr, g["freqs"])


def test_welch_matches_scipy_coherence():
    g = golden("welch.npz")
    c = oc.welch_msc(g["eeg"], g["emg"], nperseg=256)
    np.testing.assert_allclose(c, g["coh"], rtol=0, atol=1e-12)


def test_scalars():
    g = golden("scalars.npz")
    assert abs(oc.independence_threshold(5, 0.05) - float(g["IT_K5_a05"])) < 1e-15
    assert abs(oc.independence_threshold(5, 0.05) - 0.8107446225622291) < 1e-12
    assert abs(oc.independence_threshold(7, 0.01) - float(g["IT_K7_a01"])) < 1e-15
    np.testing.assert_allclose(oc.fisher_atanh_transform(g["fisher_in"]), g["fisher_out"], rtol=1e-14)
    np.testing.assert_allclose(oc.inverse_fisher_atanh(g["inv_in"]), g["inv_out"], rtol=1e-14)


def test_dpss_eigenvalues_known_answer():
    # eigenvalues for N=4096, NW=3, Kmax=5; all > 0.9 => K = 5
    _, eigs = oc.dpss_tapers(4096, 3)
    np.testing.assert_allclose(eigs, [0.99999987, 0.99999075, 0.99971499, 0.99491441, 0.946138],
                               atol=2e-6)


def test_cbpa_oracle_against_independent_scipy_implementation():
    """MNE is not installed here, so the CBPA restatement is cross-checked against a second, independent
    implementation assembled from scipy: scipy.stats.ttest_1samp for the t-map and scipy.ndimage.label
    (4-connectivity on a time x channel lattice) for the clusters, with the cluster statistics, the
    max-statistic null and MNE's p-value rule re-derived here from their definitions."""
    from scipy import ndimage, sparse, stats
    from oracle import cbpa as ocb
    rng = np.random.default_rng(2024)
    n_subj, n_times, n_ch = 11, 24, 9
    X = rng.standard_normal((n_subj, n_times, n_ch))
    X[:, 5:12, 2:6] += 1.1
    X[:, 15:20, 6:9] -= 1.2
    chain = sparse.diags([np.ones(n_ch - 1), np.ones(n_ch - 1)], [-1, 1], format="csr")    # channel c ~ c +- 1
    adj = ocb.combine_adjacency(n_times, chain)
    thr = stats.t.ppf(0.975, n_subj - 1)
    signs = np.where(rng.random((40, n_subj)) < 0.5, -1, 1).astype(np.int8)
    cross = ndimage.generate_binary_structure(2, 1)

    def independent(Xs, tail):
        t = stats.ttest_1samp(Xs, 0.0, axis=0).statistic
        tf = np.rint(np.clip(t, -ocb.T_CLAMP, ocb.T_CLAMP) * ocb.FIX_SCALE).astype(np.int64)
        masks = [t > thr, t < -thr] if tail == 0 else ([t > thr] if tail == 1 else [t < -thr])
        label_map = np.zeros(t.shape, dtype=np.int32)
        fixed, flt = [], []
        for m in masks:
            lab, n = ndimage.label(m, structure=cross)       # numbered in raster order of the first element
            for k in range(1, n + 1):
                sel = lab == k
                fixed.append(int(tf[sel].sum()))
                flt.append(float(t[sel].sum()))
                label_map[sel] = len(fixed)
        return t, label_map, np.array(fixed, dtype=np.int64), np.array(flt)

    for tail in (0, 1, -1):
        thr_call = -thr if tail == -1 else thr
        ref = ocb.permutation_cluster_1samp_test(X, signs, thr_call, tail, adj)
        t, label_map, fixed, flt = independent(X, tail)
        np.testing.assert_allclose(ref["t_obs"], t, rtol=1e-12)
        assert len(ref["clusters"]) == len(fixed) >= 1
        np.testing.assert_array_equal(ref["labels"].reshape(n_times, n_ch), label_map)     # same clusters, same order
        np.testing.assert_array_equal(ref["mass_fixed"], fixed)
        np.testing.assert_allclose(ref["mass_float"], flt, rtol=1e-12)
        h0 = np.zeros(1 + len(signs), dtype=np.int64)
        pick = {0: lambda v: v[np.argmax(np.abs(v))], 1: lambda v: v.max(), -1: lambda v: v.min()}[tail]
        h0[0] = abs(pick(fixed)) if tail == 0 else pick(fixed)
        for p, s in enumerate(signs):
            _, _, fx, _ = independent(X * s[:, None, None], tail)
            h0[1 + p] = 0 if len(fx) == 0 else fx[np.argmax(np.abs(fx))]
        np.testing.assert_array_equal(ref["H0_fixed"], h0)
        if tail == 0:
            pv = [(np.abs(h0) >= abs(v)).mean() for v in fixed]
        elif tail == 1:
            pv = [(h0 >= v).mean() for v in fixed]
        else:
            pv = [(h0 <= v).mean() for v in fixed]
        np.testing.assert_array_equal(ref["cluster_pv"], np.array(pv))
