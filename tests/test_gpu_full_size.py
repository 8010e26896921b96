"""BASELINE.json configurations at FULL size, checked through size-independent properties and through the
oracle on a channel subset (the coherence of a pair does not depend on the other channels)."""
import numpy as np
import pytest
import torch
from scipy import signal
from scipy.stats import t as t_dist

from oracle import coherence as oc
from oracle import surrogate as osur
from multimodal_biosignal_analysis_b200 import synthetic as syn

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cfg2(cuda_device):
    from multimodal_biosignal_analysis_b200 import signal_features as sf
    eeg, emg = syn.make_epochs(30, 8192, 64, 64, seed=20260102)
    emg[:, 63] = eeg[:, 5]                         # one identical pair -> coherence 1 at every frequency
    starts = syn.epoch_segment_starts(30, 8192, 2048, 1024)
    pc = sf.welch_magnitude_squared_coherence(eeg, emg, 2048.0, nperseg=2048, freq_band=(1, 100), segment_starts=starts)
    return eeg, emg, starts, pc


def test_cfg2_full_size_subset_vs_oracle_and_properties(cfg2):
    from multimodal_biosignal_analysis_b200 import signal_features as sf
    eeg, emg, starts, pc = cfg2
    coh = pc.coherence
    assert coh.shape == (100, 64, 64) and pc.n_terms == 210
    assert np.all(coh >= 0) and np.all(coh <= 1)
    assert np.max(np.abs(coh[:, 5, 63] - 1.0)) < 2e-5
    ie, im = [0, 5, 17, 40, 63], [1, 2, 31, 62, 63]
    win = signal.get_window("hann", 2048)[None]
    Xo = oc.segment_spectra(eeg[:, ie], starts, win, 1, 1, 100)[:, 0]
    Yo = oc.segment_spectra(emg[:, im], starts, win, 1, 1, 100)[:, 0]
    ref, sxx, syy, _ = oc.msc_from_spectra(Xo, Yo)
    assert np.max(np.abs(coh[:, ie][:, :, im] - ref)) < 1e-4
    np.testing.assert_allclose(pc.sxx[:, ie], sxx, rtol=5e-5)
    np.testing.assert_allclose(pc.syy[:, im], syy, rtol=5e-5)
    # scale invariance: coherence does not change when channels are rescaled (float32 exact powers of two)
    pc2 = sf.welch_magnitude_squared_coherence(eeg * 4.0, emg * 0.25, 2048.0, nperseg=2048, freq_band=(1, 100),
                                               segment_starts=starts)
    assert np.max(np.abs(pc2.coherence - coh)) < 2e-6
    # coupled pairs (common 20 / 40 Hz source) stand out at 20 Hz against the median pair
    assert coh[19].max() > 10 * np.median(coh[19])


def test_cfg3_full_size_surrogate_nulls(cfg2):
    """1,000 surrogates per pair (config 3) in both modes: counts bounded, sharding-invariant, the identical and the
    coupled pairs are significant, the null of uncoupled pairs is calibrated."""
    from multimodal_biosignal_analysis_b200 import data_surrogation as ds, kernels as K
    eeg, emg, starts, pc = cfg2
    coh = pc.coherence
    for mode in ("shift", "phase"):
        null = (ds.circular_shift_surrogate_null(pc, 1000, seed=3) if mode == "shift"
                else ds.phase_randomised_surrogate_null(pc, 1000, seed=3))
        ex = null["exceed"]
        assert ex.shape == coh.shape and ex.min() >= 0 and ex.max() <= 1000
        assert np.all(ex[:, 5, 63] == 0)                       # C = 1 is never reached by a surrogate
        p = null["p_values"]
        strong = coh > 0.2
        assert strong.sum() > 50 and np.all(p[strong] < 0.01)
        # most pairs are uncoupled (8 x 16 of 64 x 64 carry the common source): their p-values stay ~uniform
        assert 0.35 < p.mean() < 0.6 and np.mean(p < 0.05) < 0.2
        assert null["max_stat"].shape == (1000,) and 0 < null["threshold_fwe"] < 1
    # phase null: two shards with global indices reproduce the single run exactly
    res = pc.device_result
    e_all, m_all = K.surrogate_null(res, K.SURR_PHASE, 0, 1000, seed=3)
    e_a, m_a = K.surrogate_null(res, K.SURR_PHASE, 0, 373, seed=3)
    e_b, m_b = K.surrogate_null(res, K.SURR_PHASE, 373, 1000, seed=3, exceed=e_a)
    np.testing.assert_array_equal(e_b.cpu().numpy(), e_all.cpu().numpy())
    np.testing.assert_array_equal(torch.cat([m_a, m_b]).cpu().numpy(), m_all.cpu().numpy())


@pytest.mark.parametrize("mode", ["phase", "shift"])
def test_cfg3_full_size_all_surrogates_vs_fp64_on_channel_subset(cfg2, mode):
    """Config 3 at full size against the UNQUANTISED fp64 definition: the exceedance counts of ALL 1,000 surrogates
    of the 64 x 64 x 100 problem are compared, on an 8 x 8 channel subset (coupled, uncoupled and the identical
    pair), with fp64 counts - a surrogate coherence of a pair does not depend on the other channels.  Counts must
    sit inside the +-1e-4 band of the north star; cells whose count differs from the exact fp64 count (one of the
    cell's 1,000 surrogates lands within rounding of C_obs - the spectra themselves are float32 on the device) are
    reported and stay a small minority with |d count| <= 3; the per-surrogate maxima of the same subset (the kernel
    run on the subset alone) give max |dC|."""
    from multimodal_biosignal_analysis_b200 import kernels as K
    eeg, emg, starts, pc = cfg2
    res = pc.device_result
    n_surr, L = 1000, len(starts)
    coh = pc.coherence
    # the 4 most coherent EEG / EMG channels + 4 others (5 / 63 = the identical pair)
    ie = sorted(set(np.argsort(-coh.max(axis=(0, 2)))[:4].tolist() + [0, 5, 33, 62]))[:8]
    im = sorted(set(np.argsort(-coh.max(axis=(0, 1)))[:4].tolist() + [1, 31, 47, 63]))[:8]
    shifts = np.random.default_rng(3).integers(1, L, n_surr).astype(np.int32)
    kw = dict(seed=3) if mode == "phase" else dict(shifts=torch.as_tensor(shifts).cuda())
    kmode = K.SURR_PHASE if mode == "phase" else K.SURR_SHIFT
    exceed, _ = K.surrogate_null(res, kmode, 0, n_surr, **kw)
    got = exceed.cpu().numpy().astype(np.int64)[:, ie][:, :, im]
    win = signal.get_window("hann", 2048)[None]
    Xw, _ = osur.whiten(oc.segment_spectra(eeg[:, ie], starts, win, 1, 1, 100)[:, 0])
    Yw, _ = osur.whiten(oc.segment_spectra(emg[:, im], starts, win, 1, 1, 100)[:, 0])
    coh_obs = coh[:, ie][:, :, im].astype(np.float64)
    lo = np.zeros(coh_obs.shape, np.int64)
    hi = np.zeros(coh_obs.shape, np.int64)
    exact = np.zeros(coh_obs.shape, np.int64)
    ms_ref = np.zeros(n_surr)
    for s0 in range(0, n_surr, 50):                              # chunks keep the fp64 stack small
        idx = np.arange(s0, min(s0 + 50, n_surr))
        cs = osur.surrogate_coherence(Xw, Yw, mode, idx, shifts=shifts, seed=3)
        lo += (cs >= coh_obs[None] + 1e-4).sum(axis=0)
        hi += (cs >= coh_obs[None] - 1e-4).sum(axis=0)
        exact += (cs >= coh_obs[None]).sum(axis=0)
        ms_ref[idx] = cs.reshape(len(idx), -1).max(axis=1)
    assert np.all(got >= lo) and np.all(got <= hi)
    n_diff = int((got != exact).sum())
    print(f"cfg3 {mode}: {n_diff} of {got.size} cells differ from the exact fp64 count "
          f"(max |d count| {int(np.abs(got - exact).max())} of {n_surr})")
    # one near tie moves a count by the multiplicity of that surrogate: 1 for phases, the number of surrogates that
    # drew the same shift (1,000 draws from L - 1 = 209 shifts) for the shift null
    worst = 3 if mode == "phase" else 3 * int(np.bincount(shifts).max())
    assert n_diff <= got.size * 0.15 and np.abs(got - exact).max() <= worst
    # the same subset as its own 8 x 8 problem: per-surrogate max statistic vs fp64
    sub = K.csd_msc(_subset_spectra(eeg, starts, ie), _subset_spectra(emg, starts, im))
    _, ms = K.surrogate_null(sub, kmode, 0, n_surr, **kw)
    d = float(np.max(np.abs(ms.cpu().numpy() - ms_ref)))
    print(f"cfg3 {mode}: max |dC| of the per-surrogate maxima vs fp64 = {d:.2e}")
    assert d < 1e-4


def _subset_spectra(sig, starts, ch):
    from multimodal_biosignal_analysis_b200 import kernels as K
    win = torch.as_tensor(signal.get_window("hann", 2048).astype(np.float32)[None]).cuda()
    x = torch.as_tensor(np.ascontiguousarray(sig[:, ch])).cuda()
    return K.fft_segments(x, torch.as_tensor(starts).cuda(), win, 1, 1, 100)[:, 0]


def _cfg4_problem():
    from multimodal_biosignal_analysis_b200 import cbpa as cb
    X = syn.make_cbpa_contrast(20, 100, 64)
    adj = cb.combine_adjacency(100, cb.find_ch_adjacency_from_positions(syn.sensor_positions(64)))
    return X, adj, float(t_dist.ppf(0.975, 19))


def test_cfg5_cbpa_ten_thousand_permutations_vs_oracle(cuda_device):
    """config 5 CBPA count: ALL 10,000 sign-flip permutations against oracle/cbpa.py (MNE's algorithm restated,
    ~30 s of numpy / scipy): H0 bit-exact in fixed point, cluster p-values identical; plus the statistical
    properties (H0 sign-symmetric, p monotone in |mass|, the planted effect is the significant cluster)."""
    from oracle import cbpa as ocb
    from multimodal_biosignal_analysis_b200 import cbpa as cb
    X, adj, thr = _cfg4_problem()
    signs = cb.make_sign_table(10000, 20, seed=42, tail=0)
    assert signs.shape == (9999, 20)
    t_obs, clusters, pv, H0, det = cb.permutation_cluster_1samp_test(
        X, threshold=thr, n_permutations=10000, tail=0, adjacency=adj, out_type="mask", signs=signs,
        return_details=True)
    with np.errstate(all="ignore"):
        ref = ocb.permutation_cluster_1samp_test(X, signs, thr, 0, adj)
    np.testing.assert_array_equal(det["H0_fixed"], ref["H0_fixed"])
    np.testing.assert_array_equal(det["mass_fixed"], ref["mass_fixed"])
    np.testing.assert_array_equal(det["labels"].reshape(-1), ref["labels"])
    np.testing.assert_array_equal(pv, ref["cluster_pv"])
    np.testing.assert_array_equal(t_obs.reshape(-1), ref["t_obs"].reshape(-1))
    assert H0.shape == (10000,) and len(clusters) == len(pv)
    mass = np.abs(det["mass_fixed"])
    order = np.argsort(mass)
    assert np.all(np.diff(pv[order]) <= 0)                      # larger |mass| -> smaller or equal p
    best = int(np.argmax(mass))
    assert pv[best] == 1.0 / 10000 or pv[best] < 0.001
    t0 = 100 // 3
    assert clusters[best][t0:t0 + 3, :10].mean() > 0.8          # the planted block
    h = H0[1:]
    assert abs(np.mean(h > 0) - 0.5) < 0.03                     # sign symmetry of the permutation distribution
