"""GPU parity: K2 (tcgen05 pooled CSD -> MSC) and K3 (surrogate nulls) vs the fp64 oracle.
Coherence tolerance 1e-4 absolute (north star) - for the observed pass AND for every surrogate coherence, against
the UNQUANTISED float64 definition of oracle/surrogate.py (shift: 3xTF32 contraction; phase: FP16 operands with
FP32 accumulation).  Exceedance counts must lie inside the band obtained by moving the oracle's comparison by
+-SURR_TOL (counts are exact whenever no surrogate falls that close to C_obs)."""
import numpy as np
import pytest
import torch
from scipy import signal

from oracle import coherence as oc
from oracle import surrogate as osur
from multimodal_biosignal_analysis_b200 import synthetic as syn

pytestmark = pytest.mark.gpu
COH_TOL = 1e-4
SURR_TOL = 1e-4          # per-surrogate max statistic and count bands vs the unquantised fp64 definition


def _dev(a, dtype=None):
    return torch.as_tensor(np.ascontiguousarray(a), dtype=dtype).cuda()


def _welch_spectra(eeg, emg, starts, N, lo, hi):
    from multimodal_biosignal_analysis_b200 import kernels as K
    win = signal.get_window("hann", N).astype(np.float32)[None]
    X = K.fft_segments(_dev(eeg), _dev(starts), _dev(win), 1, lo, hi)[:, 0]
    Y = K.fft_segments(_dev(emg), _dev(starts), _dev(win), 1, lo, hi)[:, 0]
    return X, Y


def _oracle_spectra(eeg, emg, starts, N, lo, hi):
    win = signal.get_window("hann", N)[None]
    X = oc.segment_spectra(eeg, starts, win, 1, lo, hi)[:, 0]
    Y = oc.segment_spectra(emg, starts, win, 1, lo, hi)[:, 0]
    return X, Y


@pytest.mark.parametrize("n_epochs,ne,nm", [(6, 64, 64), (1, 11, 64), (3, 3, 70), (2, 70, 5), (5, 1, 1)])
def test_pooled_welch_coherence_matches_oracle(cuda_device, n_epochs, ne, nm):
    from multimodal_biosignal_analysis_b200 import kernels as K
    N, hop, ep = 512, 256, 2048
    eeg, emg = syn.make_epochs(n_epochs, ep, ne, nm, seed=100 + n_epochs)
    starts = syn.epoch_segment_starts(n_epochs, ep, N, hop)
    X, Y = _welch_spectra(eeg, emg, starts, N, 1, 100)
    res = K.csd_msc(X, Y, want_sxy=True)
    Xo, Yo = _oracle_spectra(eeg, emg, starts, N, 1, 100)
    coh, sxx, syy, sxy = oc.msc_from_spectra(Xo, Yo)
    assert np.max(np.abs(res.coh.cpu().numpy() - coh)) < COH_TOL
    np.testing.assert_allclose(res.sxx.cpu().numpy(), sxx, rtol=2e-5)
    np.testing.assert_allclose(res.syy.cpu().numpy(), syy, rtol=2e-5)
    scale = np.sqrt(sxx[:, :, None] * syy[:, None, :])
    assert np.max(np.abs(res.sxy.cpu().numpy() - sxy) / scale) < 1e-4


def test_pooled_coherence_tighter_than_single_tf32(cuda_device):
    """The 3xTF32 split must deliver far better than plain TF32 (~5e-5): check 2e-5 everywhere."""
    from multimodal_biosignal_analysis_b200 import kernels as K
    N, hop, ep = 512, 256, 2048
    eeg, emg = syn.make_epochs(8, ep, 64, 64, seed=7, )
    emg[:, 0] = eeg[:, 0]                      # identical pair -> C = 1
    starts = syn.epoch_segment_starts(8, ep, N, hop)
    X, Y = _welch_spectra(eeg, emg, starts, N, 1, 100)
    res = K.csd_msc(X, Y)
    Xo, Yo = _oracle_spectra(eeg, emg, starts, N, 1, 100)
    coh = oc.msc_from_spectra(Xo, Yo)[0]
    got = res.coh.cpu().numpy()
    assert np.max(np.abs(got - coh)) < 2e-5
    assert np.max(np.abs(got[:, 0, 0] - 1.0)) < 2e-5


def test_pooled_matches_scipy_coherence_cfg1(cuda_device):
    """BASELINE config 1: one EEG x one bipolar EMG channel, 60 s @ 2048 Hz, nperseg 1024."""
    from multimodal_biosignal_analysis_b200 import kernels as K
    eeg, emg = syn.make_recording(122880, 1, 2, seed=1)
    bip = (emg[:, :1] - emg[:, 1:2]).astype(np.float32)
    f, ref = signal.coherence(eeg[:, 0].astype(np.float64), bip[:, 0].astype(np.float64), fs=2048.0, nperseg=1024)
    starts = oc.welch_segments(122880, 1024)
    assert len(starts) == 239
    X, Y = _welch_spectra(eeg, bip, starts, 1024, 0, 512)
    got = K.csd_msc(X, Y).coh.cpu().numpy()[:, 0, 0]
    assert np.max(np.abs(got - ref)) < COH_TOL


def test_shift_surrogates_match_oracle(cuda_device):
    from multimodal_biosignal_analysis_b200 import kernels as K
    N, hop, ep, n_epochs = 512, 256, 2048, 6
    eeg, emg = syn.make_epochs(n_epochs, ep, 20, 70, seed=11)
    starts = syn.epoch_segment_starts(n_epochs, ep, N, hop)
    L = len(starts)
    X, Y = _welch_spectra(eeg, emg, starts, N, 1, 40)
    res = K.csd_msc(X, Y)
    rng = np.random.default_rng(3)
    n_surr = 60
    shifts = rng.integers(1, L, n_surr).astype(np.int32)
    exceed, max_stat = K.surrogate_null(res, K.SURR_SHIFT, 0, n_surr, shifts=_dev(shifts))
    Xo, Yo = _oracle_spectra(eeg, emg, starts, N, 1, 40)
    Xw, _ = osur.whiten(Xo)
    Yw, _ = osur.whiten(Yo)
    cs = osur.surrogate_coherence(Xw, Yw, "shift", np.arange(n_surr), shifts=shifts)
    coh_obs = res.coh.cpu().numpy().astype(np.float64)
    lo_cnt, _ = osur.null_statistics(cs, coh_obs, tol=+SURR_TOL)
    hi_cnt, ms = osur.null_statistics(cs, coh_obs, tol=-SURR_TOL)
    got = exceed.cpu().numpy().astype(np.int64)
    assert np.all(got >= lo_cnt) and np.all(got <= hi_cnt)
    exact, _ = osur.null_statistics(cs, coh_obs, tol=0.0)
    assert np.mean(got != exact) < 0.002                     # and almost always equal the exact count
    # 3xTF32: the surrogate coherences are as accurate as the observed pass
    assert np.max(np.abs(max_stat.cpu().numpy() - ms)) < 2e-5
    # sharding invariance: two halves accumulate to the same counts
    e2, m_a = K.surrogate_null(res, K.SURR_SHIFT, 0, 30, shifts=_dev(shifts[:30]))
    e2, m_b = K.surrogate_null(res, K.SURR_SHIFT, 30, 60, shifts=_dev(shifts[30:]), exceed=e2)
    np.testing.assert_array_equal(e2.cpu().numpy(), exceed.cpu().numpy())
    np.testing.assert_array_equal(torch.cat([m_a, m_b]).cpu().numpy(), max_stat.cpu().numpy())


def test_shift_surrogates_multitaper_groups(cuda_device):
    """windows x tapers pooling: whole windows rotate, taper index stays aligned (group = K)."""
    from multimodal_biosignal_analysis_b200 import kernels as K
    rng = np.random.default_rng(5)
    N, W, Kt = 256, 9, 5
    eeg, emg = syn.make_recording(N * (W + 1) // 2 + N, 7, 6, seed=21)
    starts = (np.arange(W) * (N // 2)).astype(np.int64)
    tapers, _ = oc.dpss_tapers(N, 3)
    X = K.fft_segments(_dev(eeg), _dev(starts), _dev(tapers, torch.float32), 0, 2, 60)
    Y = K.fft_segments(_dev(emg), _dev(starts), _dev(tapers, torch.float32), 0, 2, 60)
    F = X.shape[2]
    Xp, Yp = X.reshape(W * Kt, F, 7), Y.reshape(W * Kt, F, 6)
    res = K.csd_msc(Xp, Yp)
    Xo = oc.segment_spectra(eeg, starts, tapers, 0, 2, 60).reshape(W * Kt, F, 7)
    Yo = oc.segment_spectra(emg, starts, tapers, 0, 2, 60).reshape(W * Kt, F, 6)
    assert np.max(np.abs(res.coh.cpu().numpy() - oc.msc_from_spectra(Xo, Yo)[0])) < COH_TOL
    shifts = rng.integers(1, W, 25).astype(np.int32)
    exceed, max_stat = K.surrogate_null(res, K.SURR_SHIFT, 0, 25, shifts=_dev(shifts), group=Kt)
    Xw, _ = osur.whiten(Xo)
    Yw, _ = osur.whiten(Yo)
    cs = osur.surrogate_coherence(Xw, Yw, "shift", np.arange(25), shifts=shifts, group=Kt)
    coh_obs = res.coh.cpu().numpy().astype(np.float64)
    lo_cnt, _ = osur.null_statistics(cs, coh_obs, tol=+SURR_TOL)
    hi_cnt, ms = osur.null_statistics(cs, coh_obs, tol=-SURR_TOL)
    got = exceed.cpu().numpy().astype(np.int64)
    assert np.all(got >= lo_cnt) and np.all(got <= hi_cnt)
    assert np.max(np.abs(max_stat.cpu().numpy() - ms)) < 2e-5


def _phase_table_from_lib():
    from multimodal_biosignal_analysis_b200 import _lib
    buf = np.zeros((4096, 2), np.float32)
    assert _lib.load().cmc_phase_table(buf.ctypes.data) == 0
    return buf[:, 0] + 1j * buf[:, 1]


@pytest.mark.parametrize("ne,nm,n_surr,F_hi", [(20, 70, 150, 12), (64, 64, 130, 6), (3, 5, 40, 9)])
def test_phase_surrogates_match_oracle(cuda_device, ne, nm, n_surr, F_hi):
    """FP16 tensor-core GEMM vs the UNQUANTISED fp64 definition (exact unit-circle phases, float64
    cross-products): counts sit inside the +-1e-4 band and every per-surrogate max statistic agrees to 1e-4;
    the kernel's deviation is attributed by the FP16 emulation of the oracle (2e-6 = accumulation order).
    The surrogate index is global, so shards reproduce the unsharded run exactly."""
    from multimodal_biosignal_analysis_b200 import kernels as K
    N, hop, ep, n_epochs = 512, 256, 2048, 6
    eeg, emg = syn.make_epochs(n_epochs, ep, ne, nm, seed=31)
    starts = syn.epoch_segment_starts(n_epochs, ep, N, hop)
    X, Y = _welch_spectra(eeg, emg, starts, N, 1, F_hi)
    res = K.csd_msc(X, Y)
    seed = 0x1234ABCD5678
    exceed, max_stat = K.surrogate_null(res, K.SURR_PHASE, 0, n_surr, seed=seed)
    np.testing.assert_array_equal(_phase_table_from_lib(), osur.kernel_phase_table())
    Xo, Yo = _oracle_spectra(eeg, emg, starts, N, 1, F_hi)
    Xw, _ = osur.whiten(Xo)
    Yw, _ = osur.whiten(Yo)
    cs = osur.surrogate_coherence(Xw, Yw, "phase", np.arange(n_surr), seed=seed)      # the definition
    coh_obs = res.coh.cpu().numpy().astype(np.float64)
    lo_cnt, _ = osur.null_statistics(cs, coh_obs, tol=+SURR_TOL)
    hi_cnt, ms = osur.null_statistics(cs, coh_obs, tol=-SURR_TOL)
    got = exceed.cpu().numpy().astype(np.int64)
    assert np.all(got >= lo_cnt) and np.all(got <= hi_cnt)
    exact, _ = osur.null_statistics(cs, coh_obs, tol=0.0)
    assert np.mean(got != exact) < 0.005
    assert np.max(np.abs(max_stat.cpu().numpy() - ms)) < SURR_TOL
    # L = 42 <= 85 segments: the error-compensated operand split runs - far inside the gate
    assert np.max(np.abs(max_stat.cpu().numpy() - ms)) < 1e-5
    half = n_surr // 2
    e2, m_a = K.surrogate_null(res, K.SURR_PHASE, 0, half, seed=seed)
    e2, m_b = K.surrogate_null(res, K.SURR_PHASE, half, n_surr, seed=seed, exceed=e2)
    np.testing.assert_array_equal(e2.cpu().numpy(), exceed.cpu().numpy())
    np.testing.assert_array_equal(torch.cat([m_a, m_b]).cpu().numpy(), max_stat.cpu().numpy())


def test_phase_surrogate_null_is_calibrated(cuda_device):
    """Independent noise: p-values of the observed coherence under the phase null are ~uniform."""
    from multimodal_biosignal_analysis_b200 import kernels as K
    rng = np.random.default_rng(8)
    N, n_seg = 256, 64
    eeg = rng.standard_normal((N * n_seg, 8)).astype(np.float32)
    emg = rng.standard_normal((N * n_seg, 8)).astype(np.float32)
    starts = (np.arange(n_seg) * N).astype(np.int64)
    X, Y = _welch_spectra(eeg, emg, starts, N, 1, 64)
    res = K.csd_msc(X, Y)
    n_surr = 512
    exceed, max_stat = K.surrogate_null(res, K.SURR_PHASE, 0, n_surr, seed=11)
    p = (1.0 + exceed.cpu().numpy().astype(np.float64)) / (1.0 + n_surr)
    assert abs(p.mean() - 0.5) < 0.03
    assert abs(np.mean(p < 0.05) - 0.05) < 0.02
    assert np.all(max_stat.cpu().numpy() > 0) and np.all(max_stat.cpu().numpy() <= 1)


def test_phase_surrogates_long_segment_axis_streams_the_panel(cuda_device):
    """L = 300 segments (2L = 600 > 512): the phase panel no longer fits in shared memory and is streamed with
    the B tiles; single-term FP16 operands (L > 85), same definition, same 1e-4 gate - the deviation is the operand
    rounding the oracle can emulate - and shards still reproduce the unsharded run."""
    from multimodal_biosignal_analysis_b200 import kernels as K
    rng = np.random.default_rng(77)
    N, L, ne, nm, n_surr = 64, 300, 5, 70, 150
    t = np.arange(N * L)
    common = np.sin(2 * np.pi * 0.11 * t)
    eeg = (rng.standard_normal((N * L, ne)) + 0.5 * common[:, None]).astype(np.float32)
    emg = (rng.standard_normal((N * L, nm)) + 0.5 * np.roll(common, 2)[:, None]).astype(np.float32)
    starts = (np.arange(L) * N).astype(np.int64)
    X, Y = _welch_spectra(eeg, emg, starts, N, 2, 10)
    res = K.csd_msc(X, Y)
    seed = 99
    exceed, max_stat = K.surrogate_null(res, K.SURR_PHASE, 0, n_surr, seed=seed)
    Xo, Yo = _oracle_spectra(eeg, emg, starts, N, 2, 10)
    Xw, _ = osur.whiten(Xo)
    Yw, _ = osur.whiten(Yo)
    cs = osur.surrogate_coherence(Xw, Yw, "phase", np.arange(n_surr), seed=seed)
    coh_obs = res.coh.cpu().numpy().astype(np.float64)
    lo_cnt, _ = osur.null_statistics(cs, coh_obs, tol=+SURR_TOL)
    hi_cnt, ms = osur.null_statistics(cs, coh_obs, tol=-SURR_TOL)
    got = exceed.cpu().numpy().astype(np.int64)
    assert np.all(got >= lo_cnt) and np.all(got <= hi_cnt)
    assert np.max(np.abs(max_stat.cpu().numpy() - ms)) < SURR_TOL
    cs_emul = osur.surrogate_coherence(Xw, Yw, "phase", np.arange(40), seed=seed, emulate_fp16=True)
    assert np.max(np.abs(max_stat.cpu().numpy()[:40] - cs_emul.reshape(40, -1).max(axis=1))) < 5e-6
    e2, m_a = K.surrogate_null(res, K.SURR_PHASE, 0, 70, seed=seed)
    e2, m_b = K.surrogate_null(res, K.SURR_PHASE, 70, n_surr, seed=seed, exceed=e2)
    np.testing.assert_array_equal(e2.cpu().numpy(), exceed.cpu().numpy())
    np.testing.assert_array_equal(torch.cat([m_a, m_b]).cpu().numpy(), max_stat.cpu().numpy())


@pytest.mark.parametrize("n_epochs,ne,nm,N", [(6, 64, 64, 512), (1, 12, 64, 512), (3, 4, 70, 256), (2, 70, 6, 512),
                                                (5, 2, 2, 128), (9, 38, 20, 1024), (3, 11, 64, 256)])
def test_direct_path_matches_packed_path_and_oracle(cuda_device, n_epochs, ne, nm, N):
    """Direct K2 kernel (spectra staged as MN-major operands, 2 x 2 fold in the epilogue) against the packed-
    operand path and the fp64 oracle; odd channel pitches (11) take the packed path inside cmc_csd_coherence."""
    from multimodal_biosignal_analysis_b200 import kernels as K
    ep = 4 * N
    eeg, emg = syn.make_epochs(n_epochs, ep, ne, nm, seed=6)
    starts = syn.epoch_segment_starts(n_epochs, ep, N, N // 2)
    lo, hi = 1, min(40, N // 2)
    X, Y = _welch_spectra(eeg, emg, starts, N, lo, hi)
    ref = K.csd_msc(X, Y, want_sxy=True, keep_operands=True)
    got = K.csd_msc(X, Y, want_sxy=True)
    np.testing.assert_allclose(got.sxx.cpu().numpy(), ref.sxx.cpu().numpy(), rtol=2e-6)
    np.testing.assert_allclose(got.syy.cpu().numpy(), ref.syy.cpu().numpy(), rtol=2e-6)
    scale = np.sqrt(ref.sxx.cpu().numpy()[:, :, None] * ref.syy.cpu().numpy()[:, None, :])
    assert np.max(np.abs(got.sxy.cpu().numpy() - ref.sxy.cpu().numpy()) / scale) < 1e-6
    assert np.max(np.abs(got.coh.cpu().numpy() - ref.coh.cpu().numpy())) < 2e-6
    Xo, Yo = _oracle_spectra(eeg, emg, starts, N, lo, hi)
    assert np.max(np.abs(got.coh.cpu().numpy() - oc.msc_from_spectra(Xo, Yo)[0])) < 2e-5
    L = len(starts)
    if L > 2:
        shifts = torch.arange(1, min(L, 9), dtype=torch.int32).cuda()
        e_ref, m_ref = K.surrogate_null(ref, K.SURR_SHIFT, 0, len(shifts), shifts=shifts)
        e_got, m_got = K.surrogate_null(got, K.SURR_SHIFT, 0, len(shifts), shifts=shifts)
        assert np.max(np.abs(m_got.cpu().numpy() - m_ref.cpu().numpy())) < 2e-6


@pytest.mark.parametrize("mode", ["phase", "shift"])
def test_surrogate_null_frequency_ranges_reproduce_the_full_null(cuda_device, mode):
    """Splitting the FREQUENCY axis (how ranks share one null) gives the same counts as the unsplit call and
    the per-surrogate max statistic is the max over the pieces - for ranges that do not align with the
    groups of four bins one Philox call covers."""
    from multimodal_biosignal_analysis_b200 import kernels as K
    N, hop, ep, n_epochs, ne, nm, n_surr = 256, 128, 1024, 4, 6, 9, 96
    eeg, emg = syn.make_epochs(n_epochs, ep, ne, nm, seed=41)
    starts = syn.epoch_segment_starts(n_epochs, ep, N, hop)
    X, Y = _welch_spectra(eeg, emg, starts, N, 2, 24)            # F = 23 bins
    res = K.csd_msc(X, Y)
    F = res.coh.shape[0]
    kw = dict(seed=5) if mode == "phase" else dict(
        shifts=torch.as_tensor(np.random.default_rng(1).integers(1, len(starts), n_surr).astype(np.int32)).cuda())
    m = K.SURR_PHASE if mode == "phase" else K.SURR_SHIFT
    full_e, full_m = K.surrogate_null(res, m, 0, n_surr, **kw)
    exceed, pieces = None, []
    for f0, f1 in ((0, 5), (5, 6), (6, 6), (6, 17), (17, F)):
        exceed, mx = K.surrogate_null(res, m, 0, n_surr, exceed=exceed, f_range=(f0, f1), **kw)
        pieces.append(mx)
    np.testing.assert_array_equal(exceed.cpu().numpy(), full_e.cpu().numpy())
    np.testing.assert_array_equal(torch.stack(pieces).max(dim=0).values.cpu().numpy(), full_m.cpu().numpy())


@pytest.mark.parametrize("ne,nm,N", [(64, 64, 1024), (70, 40, 2048), (8, 8, 4096)])
def test_direct_path_many_tiles_per_cta(cuda_device, ne, nm, N):
    """Full-band spectra: 513 - 2049 bins x channel tiles = several tiles per persistent CTA, which exercises the
    accumulator double buffering and the barrier phase wrap-around of the direct kernel."""
    from multimodal_biosignal_analysis_b200 import kernels as K
    ep = 4 * N
    eeg, emg = syn.make_epochs(3, ep, ne, nm, seed=8)
    starts = syn.epoch_segment_starts(3, ep, N, N // 2)
    X, Y = _welch_spectra(eeg, emg, starts, N, 0, N // 2)
    ref = K.csd_msc(X, Y, want_sxy=True, keep_operands=True)
    got = K.csd_msc(X, Y, want_sxy=True)
    assert got.coh.shape[0] * ((ne + 63) // 64) * ((nm + 63) // 64) > 2 * 148
    np.testing.assert_allclose(got.sxx.cpu().numpy(), ref.sxx.cpu().numpy(), rtol=2e-6)
    np.testing.assert_allclose(got.syy.cpu().numpy(), ref.syy.cpu().numpy(), rtol=2e-6)
    scale = np.sqrt(ref.sxx.cpu().numpy()[:, :, None] * ref.syy.cpu().numpy()[:, None, :]) + 1e-30
    assert np.max(np.abs(got.sxy.cpu().numpy() - ref.sxy.cpu().numpy()) / scale) < 1e-6
    # bins 0 and N/2 carry (numerically) zero power after the per-segment detrend: compare away from 0/0
    ok = (ref.sxx.cpu().numpy()[:, :, None] > 1e-6) & (ref.syy.cpu().numpy()[:, None, :] > 1e-6)
    assert np.max(np.abs(got.coh.cpu().numpy() - ref.coh.cpu().numpy())[ok]) < 2e-6
    # and a second call on the same stream gives the same bits (no state left behind)
    again = K.csd_msc(X, Y)
    np.testing.assert_array_equal(again.coh.cpu().numpy(), got.coh.cpu().numpy())


@pytest.mark.parametrize("L", [1, 2, 31, 32, 33, 65])
@pytest.mark.parametrize("ne,nm,F", [(2, 2, 3), (66, 4, 5)])
def test_direct_path_segment_count_edges(cuda_device, L, ne, nm, F):
    """Segment counts around the 32-segment k-block of the direct kernel (TMA zero-fills the tail) on random
    spectra, against a float64 contraction."""
    from multimodal_biosignal_analysis_b200 import kernels as K
    g = torch.Generator(device="cuda").manual_seed(1000 + L)
    X = torch.randn((L, F, ne), dtype=torch.complex64, device="cuda", generator=g)
    Y = torch.randn((L, F, nm), dtype=torch.complex64, device="cuda", generator=g) + 0.3 * X[:, :, :1]
    got = K.csd_msc(X, Y, want_sxy=True)
    Xd, Yd = X.to(torch.complex128), Y.to(torch.complex128)
    sxy = torch.einsum("lfi,lfj->fij", Xd.conj(), Yd)
    sxx = (Xd.abs() ** 2).sum(0)
    syy = (Yd.abs() ** 2).sum(0)
    coh = (sxy.abs() ** 2 / (sxx[:, :, None] * syy[:, None, :])).clamp(0, 1)
    assert torch.max(torch.abs(got.coh.double() - coh)).item() < 2e-5
    assert torch.max(torch.abs(got.sxy.to(torch.complex128) - sxy) / torch.sqrt(sxx[:, :, None] * syy[:, None, :])).item() < 1e-5
    torch.testing.assert_close(got.sxx.double(), sxx, rtol=1e-5, atol=0)
    torch.testing.assert_close(got.syy.double(), syy, rtol=1e-5, atol=0)


@pytest.mark.parametrize("n_epochs,passes,tol", [(6, 2, 2e-4), (6, 3, 5e-6), (14, 3, 1e-4), (14, 4, 1e-4)])
def test_per_pair_null_thresholds_match_fp64_quantiles(cuda_device, n_epochs, passes, tol):
    """BASELINE config 3 "significance thresholds": the per-pair (1 - alpha) quantile of the phase-surrogate null from
    the device histograms (cmc_surrogate_null_hist with per-pair windows, cmc_hist_select, zoom passes) against
    np.quantile(..., method="higher") over the fp64 surrogate coherences of oracle/surrogate.py; windowed
    histograms and below-counters against numpy on the same stack.
    n_epochs = 6: L = 42 (error-compensated operands); n_epochs = 14: L = 98 (single FP16 term)."""
    from multimodal_biosignal_analysis_b200 import kernels as K, data_surrogation as ds
    N, hop, ep, ne, nm, n_surr, seed, alpha = 512, 256, 2048, 6, 10, 300, 21, 0.05
    eeg, emg = syn.make_epochs(n_epochs, ep, ne, nm, seed=77)
    emg[:, 0] += 0.7 * eeg[:, 0]                                     # one strongly coupled pair (large |S| terms)
    emg[:, 9] = 0.0                                                  # and a silent channel: every C_s is 0
    starts = syn.epoch_segment_starts(n_epochs, ep, N, hop)
    X, Y = _welch_spectra(eeg, emg, starts, N, 2, 13)
    res = K.csd_msc(X, Y)
    Xo, Yo = _oracle_spectra(eeg, emg, starts, N, 2, 13)
    Xw, _ = osur.whiten(Xo)
    Yw, _ = osur.whiten(Yo)
    cs = osur.surrogate_coherence(Xw, Yw, "phase", np.arange(n_surr), seed=seed)          # (S, F, Ne, Nm) fp64
    want = np.quantile(cs, 1.0 - alpha, axis=0, method="higher")
    # the analytic mean of the null that places the first window
    np.testing.assert_allclose(res.null_mean().cpu().numpy(), np.einsum("lfi,lfj->fij", np.abs(Xw) ** 2, np.abs(Yw) ** 2),
                               rtol=1e-4, atol=1e-9)
    assert abs(cs[:, :, :, :9].mean() / res.null_mean().cpu().numpy()[:, :, :9].mean() - 1.0) < 0.05
    thr, last = ds.null_quantile_thresholds(res, n_surr, seed, 1.0 - alpha, passes=passes, n_bins=128)
    got = thr.cpu().numpy()
    err = np.abs(got - want)
    worst = np.unravel_index(np.argmax(err), err.shape)
    assert err.max() < tol, f"max threshold error {err.max():.2e} at {worst}: got {got[worst]:.6f} want {want[worst]:.6f}"
    assert np.all(got[:, :, 9] < 1e-4)                               # silent channel: the null is identically 0
    # windowed histograms: 64 bins over [0.01, 0.06), below-counter for everything under 0.01
    lo_t = torch.full(res.coh.shape, 0.01, device="cuda")
    sc_t = torch.full(res.coh.shape, 64 / 0.05, device="cuda")
    hist, below = K.surrogate_null_hist(res, 0, n_surr, seed=seed, n_bins=64, bin_lo=lo_t, bin_scale=sc_t)
    h, bl = hist.cpu().numpy().astype(np.int64), below.cpu().numpy().astype(np.int64)
    xb = np.floor((cs - 0.01) * (64 / 0.05)).astype(np.int64)
    ref_below = (xb < 0).sum(axis=0)
    ref = np.zeros_like(h)
    f_i, e_i, m_i = np.meshgrid(*[np.arange(k) for k in cs.shape[1:]], indexing="ij")
    for s in range(n_surr):
        inside = (xb[s] >= 0) & (xb[s] < 64)
        np.add.at(ref, (f_i[inside], e_i[inside], m_i[inside], xb[s][inside]), 1)
    # only values within rounding of a bin edge (bins are 7.8e-4 wide; the single-term FP16 operands of L = 98 move a
    # surrogate coherence by up to ~4e-5, the split operands of L = 42 by ~1e-6) may land in the neighbouring bin
    assert np.abs(h - ref).sum() <= (0.004 if n_epochs == 6 else 0.02) * n_surr * ref[..., 0].size
    assert np.abs(bl - ref_below).max() <= 4 and np.abs((bl + h.sum(-1)) - (ref_below + ref.sum(-1))).max() <= 4
    # chunks of the surrogate range accumulate to the same histogram
    h2, b2 = K.surrogate_null_hist(res, 0, 170, seed=seed, n_bins=64, bin_lo=lo_t, bin_scale=sc_t)
    h2, b2 = K.surrogate_null_hist(res, 170, n_surr, seed=seed, n_bins=64, bin_lo=lo_t, bin_scale=sc_t, hist=h2, below=b2)
    np.testing.assert_array_equal(h2.cpu().numpy(), hist.cpu().numpy())
    np.testing.assert_array_equal(b2.cpu().numpy(), below.cpu().numpy())
    # public API: thresholds + significance mask next to the exceedance p-values, histograms on request
    import types
    pooled = types.SimpleNamespace(device_result=res, coherence=res.coh, freqs=np.arange(res.coh.shape[0]))
    out = ds.phase_randomised_surrogate_null(pooled, n_surr, seed=seed, alpha=alpha, thresholds=True,
                                             threshold_passes=passes, return_hist=True, hist_bins=64,
                                             hist_range=(0.01, 0.06))
    np.testing.assert_allclose(out["threshold"], got, rtol=0, atol=1e-7)
    np.testing.assert_array_equal(out["null_hist"], hist.cpu().numpy())
    np.testing.assert_array_equal(out["null_hist_below"], below.cpu().numpy())
    coh = res.coh.cpu().numpy()
    np.testing.assert_array_equal(out["significant"], coh > out["threshold"])
    # a pair above its threshold has few exceedances, and vice versa (alpha n = 15 of 300; ties at the edge excluded)
    clear = np.abs(coh - out["threshold"]) > 1e-3
    assert np.all((out["exceed"][clear & out["significant"]] <= 15))
    assert np.all((out["exceed"][clear & ~out["significant"]] >= 15))
    assert out["significant"][:, 0, 0].all()
