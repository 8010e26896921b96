"""GPU parity: K1 (segment spectra) and K2w (per-window MSC / jackknife) vs the fp64 oracle
and the reference-generated golden vectors.  Tolerances: coherence and CI bounds 1e-4
absolute (north star); spectra 2e-6 of the segment's spectral scale."""
import numpy as np
import pytest
import torch
from scipy import signal
from scipy.stats import t as t_dist

from conftest import golden
from oracle import coherence as oc

pytestmark = pytest.mark.gpu

COH_TOL = 1e-4


def _dev(a, dtype=None):
    return torch.as_tensor(np.ascontiguousarray(a), dtype=dtype).cuda()


@pytest.mark.parametrize("N", [128, 256, 512, 1024, 2048, 4096, 8192])
@pytest.mark.parametrize("detrend", [0, 1, 2])
def test_fft_segments_matches_numpy(cuda_device, N, detrend):
    from multimodal_biosignal_analysis_b200 import kernels as K
    rng = np.random.default_rng(N + detrend)
    n_ch = 11 if N >= 2048 else 37
    n = N * 3 + 17
    x = (rng.standard_normal((n, n_ch)) + 0.7).astype(np.float32)
    starts = np.array([0, 5, N // 2 + 3, n - N], dtype=np.int64)
    wins = np.stack([signal.get_window("hann", N), signal.windows.dpss(N, 3, 2)[1]]).astype(np.float32)
    ref = oc.segment_spectra(x.astype(np.float64), starts, wins.astype(np.float64), detrend)
    got = K.fft_segments(_dev(x), _dev(starts), _dev(wins), detrend).cpu().numpy()
    assert got.shape == ref.shape
    scale = np.sqrt(np.mean(np.abs(ref) ** 2))
    assert np.max(np.abs(got - ref)) < 2e-6 * scale * np.sqrt(N)


@pytest.mark.parametrize("N", [512, 1024, 2048, 4096])
@pytest.mark.parametrize("n_ch,detrend", [(64, 1), (24, 0), (12, 2)])
def test_fft_segments_tma_paths(cuda_device, N, n_ch, detrend):
    """channel pitch multiple of 4 floats -> TMA-staged two-channel kernel; three tapers per segment, partial
    last channel tile."""
    from multimodal_biosignal_analysis_b200 import kernels as K
    rng = np.random.default_rng(N + n_ch)
    n = N * 2 + 64
    x = (rng.standard_normal((n, n_ch)) * np.linspace(0.5, 2.0, n_ch) + 0.4).astype(np.float32)
    starts = np.array([0, 37, n - N], dtype=np.int64)
    wins = np.concatenate([signal.get_window("hann", N)[None], signal.windows.dpss(N, 3, 2)]).astype(np.float32)
    ref = oc.segment_spectra(x.astype(np.float64), starts, wins.astype(np.float64), detrend)
    got = K.fft_segments(_dev(x), _dev(starts), _dev(wins), detrend).cpu().numpy()
    scale = np.sqrt(np.mean(np.abs(ref) ** 2))
    assert np.max(np.abs(got - ref)) < 2e-6 * scale * np.sqrt(N)
    lo, hi = 3, min(100, N // 2)
    # a band below the last pass's stride takes the fused last-pass + split path: same values, other rounding
    got_b = K.fft_segments(_dev(x), _dev(starts), _dev(wins), detrend, lo, hi).cpu().numpy()
    assert np.max(np.abs(got_b - ref[:, :, lo:hi + 1])) < 2e-6 * scale * np.sqrt(N)
    assert np.max(np.abs(got_b - got[:, :, lo:hi + 1])) < 1e-6 * scale * np.sqrt(N)


@pytest.mark.parametrize("N,lo,hi", [(2048, 0, 127), (2048, 1, 100), (2048, 100, 128), (2048, 127, 127), (4096, 0, 255),
                                     (4096, 17, 256), (1024, 0, 63), (1024, 5, 64), (512, 0, 15), (256, 0, 15),
                                     (128, 0, 7), (8192, 0, 255), (8192, 200, 256)])
@pytest.mark.parametrize("n_ch,detrend", [(16, 1), (12, 2), (6, 0)])
def test_fft_segments_band_limited_fused_path(cuda_device, N, lo, hi, n_ch, detrend):
    """bands that end at / just past the stride of the last radix pass (fused last pass + split vs the general
    path), with bin 0, a partial channel tile and two tapers."""
    from multimodal_biosignal_analysis_b200 import kernels as K
    rng = np.random.default_rng(N + lo + hi + n_ch)
    n = N * 2 + 40
    x = (rng.standard_normal((n, n_ch)) * np.linspace(0.5, 2.0, n_ch) - 0.3).astype(np.float32)
    starts = np.array([0, 23, n - N], dtype=np.int64)
    wins = np.stack([signal.get_window("hann", N), signal.windows.dpss(N, 3, 2)[1]]).astype(np.float32)
    ref = oc.segment_spectra(x.astype(np.float64), starts, wins.astype(np.float64), detrend)
    got = K.fft_segments(_dev(x), _dev(starts), _dev(wins), detrend, lo, hi).cpu().numpy()
    assert got.shape == (3, 2, hi - lo + 1, n_ch)
    scale = np.sqrt(np.mean(np.abs(ref) ** 2))
    assert np.max(np.abs(got - ref[:, :, lo:hi + 1])) < 2e-6 * scale * np.sqrt(N)
    if lo == 0:
        assert np.all(got[:, :, 0].imag == 0)


def test_fft_segments_band_and_offsets(cuda_device):
    from multimodal_biosignal_analysis_b200 import kernels as K
    rng = np.random.default_rng(3)
    N = 2048
    eeg = rng.standard_normal((3 * N, 64)).astype(np.float32)
    emg = rng.standard_normal((3 * N, 64)).astype(np.float32)
    starts = np.arange(0, 2 * N + 1, N // 2, dtype=np.int64)
    win = signal.get_window("hann", N).astype(np.float32)[None]
    out = torch.zeros((len(starts), 1, 100, 128), dtype=torch.complex64, device="cuda")
    K.fft_segments(_dev(eeg), _dev(starts), _dev(win), 1, 1, 100, out=out, ch_offset=0)
    K.fft_segments(_dev(emg), _dev(starts), _dev(win), 1, 1, 100, out=out, ch_offset=64)
    ref_e = oc.segment_spectra(eeg, starts, win, 1, 1, 100)
    ref_m = oc.segment_spectra(emg, starts, win, 1, 1, 100)
    got = out.cpu().numpy()
    scale = np.sqrt(np.mean(np.abs(ref_e) ** 2))
    assert np.max(np.abs(got[..., :64] - ref_e)) < 1e-4 * scale
    assert np.max(np.abs(got[..., 64:] - ref_m)) < 1e-4 * scale


def test_fft_rejects_bad_arguments(cuda_device):
    from multimodal_biosignal_analysis_b200 import kernels as K
    from multimodal_biosignal_analysis_b200._lib import CmcError
    x = torch.zeros((4000, 4), device="cuda")
    st = torch.zeros(1, dtype=torch.int64, device="cuda")
    with pytest.raises(CmcError):
        K.fft_segments(x, st, torch.ones((1, 1000), device="cuda"), 0, 0, 600)   # bin beyond N / 2
    with pytest.raises(ValueError):
        K.check_segments([3000], 2048, 4000)                                   # segment past the end
    with pytest.raises(TypeError):
        K.fft_segments(x.cpu(), st, torch.ones((1, 1024), device="cuda"))    # no CPU path


def _spectra(x, fs, win_sec, overlap, tapers):
    from multimodal_biosignal_analysis_b200 import kernels as K
    N, hop = oc.window_params(fs, win_sec, overlap)
    W = oc.n_windows_msc(x.shape[0], N, hop)
    starts = np.arange(W, dtype=np.int64) * hop
    return K.fft_segments(_dev(x, torch.float32), _dev(starts), _dev(tapers, torch.float32), 0)


def test_msc_windows_nojk_matches_reference_golden(cuda_device):
    from multimodal_biosignal_analysis_b200 import kernels as K
    g = golden("msc_nojk.npz")
    fs = float(g["fs"])
    tapers, _ = oc.dpss_tapers(256, 3)
    X = _spectra(g["eeg"], fs, 1.0, 0.5, tapers)
    Y = _spectra(g["emg"], fs, 1.0, 0.5, tapers)
    coh, lo, hi, sig = K.msc_windows(X, Y, None, False, 0.0, float(g["IT"]))
    coh = coh.cpu().numpy()
    assert lo is None and hi is None
    assert coh.shape == g["coherence_raw"].shape
    assert np.max(np.abs(coh - g["coherence_raw"])) < COH_TOL
    diff = sig.cpu().numpy().astype(bool) != g["coherence_significant"]
    assert np.all(np.abs(g["coherence_raw"][diff] - float(g["IT"])) <= COH_TOL)


def test_msc_windows_jackknife_matches_reference_golden(cuda_device):
    from multimodal_biosignal_analysis_b200 import kernels as K
    g = golden("msc_jk.npz")
    fs = float(g["fs"])
    tapers, _ = oc.dpss_tapers(256, 3)
    X = _spectra(g["eeg"], fs, 1.0, 0.5, tapers)
    Y = _spectra(g["emg"], fs, 1.0, 0.5, tapers)
    mask = _dev(g["window_mask"].astype(np.uint8))
    t_crit = t_dist.ppf(1 - 0.05 / 2, 4)
    coh, lo, hi, sig = K.msc_windows(X, Y, mask, True, t_crit, float(g["IT_bonferroni"]))
    for got, key in ((coh, "coherence_raw"), (lo, "ci_lower"), (hi, "ci_upper")):
        assert np.max(np.abs(got.cpu().numpy() - g[key])) < COH_TOL, key
    assert np.all(coh.cpu().numpy()[~g["window_mask"]] == 0)
    diff = sig.cpu().numpy().astype(bool) != g["coherence_significant"]
    assert np.all(np.abs(g["coherence_raw"][diff] - float(g["IT_bonferroni"])) <= COH_TOL)
    # fused EMG-argmax variant against the unfused result
    c2, l2, h2, arg = K.msc_windows_maxemg(X, Y, mask, True, t_crit, None, False, True)
    a, b, d = oc.max_over_emg(coh.cpu().numpy(), lo.cpu().numpy(), hi.cpu().numpy())
    np.testing.assert_array_equal(c2.cpu().numpy(), a)
    np.testing.assert_array_equal(l2.cpu().numpy(), b)
    np.testing.assert_array_equal(h2.cpu().numpy(), d)
    act = g["window_mask"]
    np.testing.assert_array_equal(arg.cpu().numpy()[act], np.argmax(coh.cpu().numpy(), axis=3)[act])


@pytest.mark.parametrize("K_tapers,nw", [(7, 4), (3, 2)])
def test_msc_windows_other_taper_counts_vs_oracle(cuda_device, K_tapers, nw):
    from multimodal_biosignal_analysis_b200 import kernels as K
    rng = np.random.default_rng(K_tapers)
    fs, N = 512.0, 512
    n = N * 4
    src = rng.standard_normal(n)
    eeg = (rng.standard_normal((n, 5)) + src[:, None] * np.linspace(0, 2, 5)).astype(np.float32)
    emg = (rng.standard_normal((n, 9)) + src[:, None] * np.linspace(2, 0, 9)).astype(np.float32)
    ref = oc.multitaper_msc(eeg, emg, fs, nw=nw, use_jackknife=True, jackknife_alpha=0.1,
                            apply_independence_threshold=False)
    tapers, _ = oc.dpss_tapers(N, nw)
    assert len(tapers) == K_tapers
    X = _spectra(eeg, fs, 1.0, 0.5, tapers)
    Y = _spectra(emg, fs, 1.0, 0.5, tapers)
    coh, lo, hi, _ = K.msc_windows(X, Y, None, True, t_dist.ppf(0.95, K_tapers - 1), None)
    assert np.max(np.abs(coh.cpu().numpy() - ref["coherence_raw"])) < COH_TOL
    # With K = 3 every leave-one-out estimate averages two tapers and sits within ~1e-5 of 1 for coupled
    # pairs; z = atanh-like then amplifies float32 rounding of the spectra (the reference's own float32
    # jackknife, signal_features.py:503-510, has the same conditioning), so the CI gate is wider there.
    ci_tol = COH_TOL if K_tapers >= 5 else 2e-3
    assert np.max(np.abs(lo.cpu().numpy() - ref["coherence_ci_lower"])) < ci_tol
    assert np.max(np.abs(hi.cpu().numpy() - ref["coherence_ci_upper"])) < ci_tol


def test_msc_identical_and_zero_channels(cuda_device):
    """identical signals -> C = 1; an all-zero (bad) channel -> C = 0, never NaN."""
    from multimodal_biosignal_analysis_b200 import kernels as K
    rng = np.random.default_rng(5)
    n, N = 2048, 512
    eeg = rng.standard_normal((n, 3)).astype(np.float32)
    emg = eeg.copy()
    emg[:, 2] = 0.0
    tapers, _ = oc.dpss_tapers(N, 3)
    X = _spectra(eeg, 512.0, 1.0, 0.5, tapers)
    Y = _spectra(emg, 512.0, 1.0, 0.5, tapers)
    coh, lo, hi, _ = K.msc_windows(X, Y, None, True, 2.776, None)
    c = coh.cpu().numpy()
    assert np.all(np.isfinite(c)) and np.all(np.isfinite(lo.cpu().numpy())) and np.all(np.isfinite(hi.cpu().numpy()))
    assert np.max(np.abs(c[:, :, 0, 0] - 1.0)) < 1e-5
    assert np.max(np.abs(c[:, :, 1, 1] - 1.0)) < 1e-5
    assert np.all(c[:, :, :, 2] == 0)


@pytest.mark.parametrize("N", [100, 1000, 1536])
def test_fft_segments_direct_dft_for_other_lengths(cuda_device, N):
    from multimodal_biosignal_analysis_b200 import kernels as K
    rng = np.random.default_rng(N)
    x = (rng.standard_normal((2 * N + 10, 5)) + 0.2).astype(np.float32)
    starts = np.array([0, N + 3], dtype=np.int64)
    win = signal.get_window("hann", N).astype(np.float32)[None]
    for detrend in (0, 1, 2):
        ref = oc.segment_spectra(x.astype(np.float64), starts, win.astype(np.float64), detrend)
        got = K.fft_segments(_dev(x), _dev(starts), _dev(win), detrend).cpu().numpy()
        scale = np.sqrt(np.mean(np.abs(ref) ** 2))
        assert np.max(np.abs(got - ref)) < 1e-5 * scale * np.sqrt(N)


@pytest.mark.parametrize("N", [512, 1024, 2048])
def test_fft_segments_pipelined_kernel_many_tiles(cuda_device, N):
    """>= 64 (segment, channel-tile) units -> the persistent two-worker / three-buffer kernel, with several window
    rows per segment (tile re-fetch path) and more tiles than workers (prefetch hand-over path)."""
    from multimodal_biosignal_analysis_b200 import kernels as K
    rng = np.random.default_rng(N)
    n_seg, n_ch = 500, 20                                   # 1500 tiles over 296 workers
    hop = N // 4
    n = hop * (n_seg - 1) + N
    x = (rng.standard_normal((n, n_ch)) + 0.1).astype(np.float32)
    starts = (np.arange(n_seg) * hop).astype(np.int64)
    wins = signal.windows.dpss(N, 3, 3).astype(np.float32)
    for detrend, nw in ((1, 1), (0, 3)):
        ref = oc.segment_spectra(x.astype(np.float64), starts, wins[:nw].astype(np.float64), detrend, 2, 90)
        got = K.fft_segments(_dev(x), _dev(starts), _dev(wins[:nw]), detrend, 2, 90).cpu().numpy()
        scale = np.sqrt(np.mean(np.abs(ref) ** 2))
        assert np.max(np.abs(got - ref)) < 2e-6 * scale * np.sqrt(N)
    # determinism: repeated launches give bit-identical spectra
    a = K.fft_segments(_dev(x), _dev(starts), _dev(wins[:1]), 1, 2, 90)
    b = K.fft_segments(_dev(x), _dev(starts), _dev(wins[:1]), 1, 2, 90)
    assert torch.equal(torch.view_as_real(a), torch.view_as_real(b))


@pytest.mark.parametrize("N,ne,nm,n_seg,n_win", [(2048, 64, 64, 12, 1), (512, 11, 70, 9, 1), (1024, 8, 5, 7, 3),
                                                   (512, 3, 3, 2, 1), (4096, 8, 8, 3, 1), (256, 6, 9, 4, 1)])
def test_fft_segments_pair_is_bit_identical_to_two_launches(cuda_device, N, ne, nm, n_seg, n_win):
    """cmc_fft_segments_pair: EEG and EMG in ONE launch of the pipelined K1 kernel (tiles of both recordings share
    the persistent CTAs) - and its fall-back to two launches for sizes / layouts outside that kernel - give exactly
    the spectra of two cmc_fft_segments calls, written into channel ranges of one array."""
    import torch
    from multimodal_biosignal_analysis_b200 import kernels as K
    rng = np.random.default_rng(N + ne)
    n = N * (n_seg + 1) // 2 + N
    x1 = torch.as_tensor(rng.standard_normal((n, ne)).astype(np.float32) + 3.0).cuda()
    x2 = torch.as_tensor(rng.standard_normal((n, nm)).astype(np.float32) - 1.0).cuda()
    starts = torch.as_tensor((np.arange(n_seg) * (N // 2)).astype(np.int64)).cuda()
    win = torch.as_tensor(rng.random((n_win, N)).astype(np.float32)).cuda()
    for lo, hi in ((1, min(100, N // 2)), (0, N // 2)):
        F = hi - lo + 1
        ne_p, nm_p = ne + (ne & 1), nm + (nm & 1)
        joint = torch.zeros((n_seg, n_win, F, ne_p + nm_p), dtype=torch.complex64, device="cuda")
        K.fft_segments_pair(x1, x2, starts, win, 1, lo, hi, joint[..., :ne], joint[..., ne_p:ne_p + nm])
        a = K.fft_segments(x1, starts, win, 1, lo, hi)
        b = K.fft_segments(x2, starts, win, 1, lo, hi)
        assert torch.equal(joint[..., :ne], a) and torch.equal(joint[..., ne_p:ne_p + nm], b)
        if ne & 1:
            assert torch.all(joint[..., ne] == 0)                  # padding columns are never written


def test_first_use_table_inside_capture_is_refused_cleanly(cuda_device):
    """The twiddle table of a new N cannot be created while the stream is captured (cudaMalloc + synchronous copy):
    the call must fail with a message that names cmc_fft_prepare, and work after preparing."""
    from multimodal_biosignal_analysis_b200 import _lib, kernels as K
    N = 160                                      # direct-DFT length that no other test uses
    x = torch.randn(4 * N, 8, device="cuda")
    starts = torch.tensor([0, N], dtype=torch.int64, device="cuda")
    win = torch.ones(1, N, device="cuda")
    out = torch.empty((2, 1, N // 2 + 1, 8), dtype=torch.complex64, device="cuda")
    g = torch.cuda.CUDAGraph()
    with pytest.raises(_lib.CmcError, match="cmc_fft_prepare"):
        with torch.cuda.graph(g):
            K.fft_segments(x, starts, win, 0, out=out)
    torch.cuda.synchronize()
    assert _lib.load().cmc_fft_prepare(N) == 0
    g2 = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g2):
        K.fft_segments(x, starts, win, 0, out=out)
    g2.replay()
    torch.cuda.synchronize()
    ref = torch.fft.rfft(x[:N].double(), dim=0)
    assert torch.allclose(out[0, 0].to(torch.complex128), ref, atol=1e-3)


def test_msc_windows_large_staging_and_unaligned_outputs(cuda_device):
    """K (Ne + Nm) large enough for more than 48 KB of staged operands (opt-in shared memory), and outputs that are
    only 4-byte aligned (scalar stores instead of the 8-byte ones): same bits as the aligned call, oracle parity."""
    from multimodal_biosignal_analysis_b200 import kernels as K, _lib
    rng = np.random.default_rng(11)
    W, Kt, F, Ne, Nm = 2, 7, 3, 128, 256
    X = (rng.standard_normal((W, Kt, F, Ne)) + 1j * rng.standard_normal((W, Kt, F, Ne))).astype(np.complex64)
    Y = (rng.standard_normal((W, Kt, F, Nm)) + 1j * rng.standard_normal((W, Kt, F, Nm))).astype(np.complex64)
    Y = (Y + 0.7 * X[..., :1]).astype(np.complex64)
    Xd, Yd = _dev(X), _dev(Y)
    t_crit = float(t_dist.ppf(0.975, Kt - 1))
    coh, lo, hi, sig = K.msc_windows(Xd, Yd, None, True, t_crit, 0.3)
    for w in range(W):
        m, l, h = oc.jackknife_from_spectra(X[w].astype(np.complex128), Y[w].astype(np.complex128), 0.05)
        assert np.max(np.abs(coh[w].cpu().numpy() - m)) < COH_TOL
        tame = m < 0.999
        assert np.max(np.abs(lo[w].cpu().numpy() - l)[tame]) < 2e-3
        assert np.max(np.abs(hi[w].cpu().numpy() - h)[tame]) < 2e-3
    best, blo, bhi, arg = K.msc_windows_maxemg(Xd, Yd, None, True, t_crit, None, return_argmax=True)
    np.testing.assert_array_equal(best.cpu().numpy(), coh.cpu().numpy().max(axis=3))
    np.testing.assert_array_equal(arg.cpu().numpy(), coh.cpu().numpy().argmax(axis=3))
    # the same call through the C ABI with every output shifted by one element
    n = W * F * Ne * Nm
    bufs = [torch.zeros(n + 1, dtype=torch.float32, device="cuda") for _ in range(3)]
    sbuf = torch.zeros(n + 1, dtype=torch.uint8, device="cuda")
    lib = _lib.load()
    rc = lib.cmc_msc_windows(Xd.data_ptr(), Yd.data_ptr(), W, Kt, F, Ne, Nm, Ne, Nm, None, 1, t_crit, 0.3,
                             bufs[0].data_ptr() + 4, bufs[1].data_ptr() + 4, bufs[2].data_ptr() + 4,
                             sbuf.data_ptr() + 1, _lib.current_stream())
    _lib.check(rc, "cmc_msc_windows")
    torch.cuda.synchronize()
    for got, ref in zip(bufs, (coh, lo, hi)):
        assert torch.equal(got[1:], ref.reshape(-1))
    assert torch.equal(sbuf[1:], sig.reshape(-1))
