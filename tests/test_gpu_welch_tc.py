"""GPU parity of the tensor-core Welch kernel (cmc_welch_hann_*: half-block BF16x3 tcgen05 DFT, hann as a three-tap
epilogue) through the C ABI against the fp64 oracle.

Gates: spectra within 1.2e-4 of the channel's spectral rms (operands carry 16 significand bits; the FFT kernel sits
at 2e-6), coherence computed from them within 1e-5 of the fp64 coherence (north star: 1e-4), bit-identical results
from run to run and from a CUDA-graph replay, and every detrend convention / band edge / channel layout the FFT
kernel takes.  What scipy.signal.coherence does per pair (preprocessing.py:1228-1230) is the ground truth."""
import numpy as np
import pytest
import torch
from scipy import signal

from oracle import coherence as oc

pytestmark = pytest.mark.gpu

SPEC_TOL = 1.2e-4


def _dev(a, dtype=None):
    return torch.as_tensor(np.ascontiguousarray(a), dtype=dtype).cuda()


def _oracle_spectra(x, starts, N, detrend, lo, hi):
    win = signal.get_window("hann", N)[None]
    return oc.segment_spectra(x.astype(np.float64), np.asarray(starts), win, detrend, lo, hi)[:, 0]


def _run(K, plan, eeg, emg, F, detrend=1):
    ne = eeg.shape[1]
    nm = 0 if emg is None else emg.shape[1]
    spec = torch.full((plan.n_seg, 1, F, ne + nm), float("nan"), dtype=torch.complex64, device="cuda")
    if emg is None:
        plan.spectra(_dev(eeg), spec, detrend=detrend)
    else:
        plan.spectra(_dev(eeg), spec[..., :ne], _dev(emg), spec[..., ne:], detrend=detrend)
    torch.cuda.synchronize()
    return spec


def _rel_err(got, ref):
    rms = np.sqrt(np.mean(np.abs(ref) ** 2, axis=(0, 1), keepdims=True)) + 1e-30
    return float(np.max(np.abs(got - ref) / rms))


@pytest.mark.parametrize("detrend,lo,hi", [(1, 1, 100), (0, 1, 100), (2, 0, 99), (1, 0, 101), (1, 40, 141), (0, 0, 50)])
def test_spectra_match_oracle_config2_epochs(cuda_device, detrend, lo, hi):
    from multimodal_biosignal_analysis_b200 import kernels as K, synthetic as syn
    eeg, emg = syn.make_epochs(3, 8192, 64, 64, seed=1)
    starts = syn.epoch_segment_starts(3, 8192, 2048, 1024)
    plan = K.WelchHannPlan(starts, 2048, lo, hi)
    assert plan.n_half_blocks == 24                       # 8 half blocks per epoch instead of 2 x 7
    got = _run(K, plan, eeg, emg, hi - lo + 1, detrend)[:, 0].cpu().numpy()
    ref = _oracle_spectra(np.concatenate([eeg, emg], axis=1), starts, 2048, detrend, lo, hi)
    assert not np.isnan(got.view(np.float32)).any()
    assert _rel_err(got, ref) < SPEC_TOL


def test_mixed_chains_odd_channel_counts(cuda_device):
    """Isolated segments, short chains, unsorted starts, 12 + 60 channels (rows of the unit left empty)."""
    from multimodal_biosignal_analysis_b200 import kernels as K
    rng = np.random.default_rng(0)
    e = rng.standard_normal((20000, 12)).astype(np.float32)
    m = rng.standard_normal((20000, 60)).astype(np.float32)
    starts = np.array([0, 512, 1024, 5000, 5512, 9000, 3000, 3512, 4024, 18976], dtype=np.int64)
    plan = K.WelchHannPlan(starts, 1024, 1, 50)
    assert plan.n_half_blocks == 15
    got = _run(K, plan, e, m, 50)[:, 0].cpu().numpy()
    ref = _oracle_spectra(np.concatenate([e, m], axis=1), starts, 1024, 1, 1, 50)
    assert _rel_err(got, ref) < SPEC_TOL


@pytest.mark.parametrize("n_ch,N,lo,hi", [(128, 512, 2, 60), (200, 4096, 1, 100), (64, 256, 1, 20)])
def test_single_recording(cuda_device, n_ch, N, lo, hi):
    from multimodal_biosignal_analysis_b200 import kernels as K
    rng = np.random.default_rng(n_ch)
    x = rng.standard_normal((3 * N + 1000, n_ch)).astype(np.float32)
    starts = np.arange(0, x.shape[0] - N, N // 2, dtype=np.int64)
    plan = K.WelchHannPlan(starts, N, lo, hi)
    got = _run(K, plan, x, None, hi - lo + 1)[:, 0].cpu().numpy()
    ref = _oracle_spectra(x, starts, N, 1, lo, hi)
    assert _rel_err(got, ref) < SPEC_TOL


def test_dc_offset_and_drift_no_worse_than_fft(cuda_device):
    """A DC level of 1000 sigma plus a ramp: float32 input quantisation bounds both kernels; the tensor-core path
    (chain offset removed exactly before the BF16 split) must stay at the FFT kernel's error level."""
    from multimodal_biosignal_analysis_b200 import kernels as K, synthetic as syn
    eeg, emg = syn.make_epochs(3, 8192, 64, 64, seed=1)
    starts = syn.epoch_segment_starts(3, 8192, 2048, 1024)
    t = np.arange(eeg.shape[0], dtype=np.float32)[:, None]
    eeg = (eeg + 1000.0 + 0.01 * t).astype(np.float32)
    emg = (emg - 300.0).astype(np.float32)
    plan = K.WelchHannPlan(starts, 2048, 1, 100)
    got = _run(K, plan, eeg, emg, 100)[:, 0].cpu().numpy()
    ref = _oracle_spectra(np.concatenate([eeg, emg], axis=1), starts, 2048, 1, 1, 100)
    win = _dev(signal.get_window("hann", 2048).astype(np.float32)[None])
    fft = torch.empty((len(starts), 1, 100, 128), dtype=torch.complex64, device="cuda")
    K.fft_segments_pair(_dev(eeg), _dev(emg), _dev(starts), win, 1, 1, 100, fft[..., :64], fft[..., 64:])
    e_tc, e_fft = _rel_err(got, ref), _rel_err(fft[:, 0].cpu().numpy(), ref)
    assert e_tc < max(2.0 * e_fft, SPEC_TOL)


def test_config2_full_size_coherence_and_determinism(cuda_device):
    """BASELINE config 2 at full size: coherence from the tensor-core spectra vs fp64, 1e-5; two launches and a
    CUDA-graph replay are bit-identical (store-then-add emission is order independent)."""
    from multimodal_biosignal_analysis_b200 import kernels as K, synthetic as syn
    eeg, emg = syn.make_epochs(30, 8192, 64, 64, seed=20260102)
    starts = syn.epoch_segment_starts(30, 8192, 2048, 1024)
    plan = K.WelchHannPlan(starts, 2048, 1, 100)
    assert plan.n_half_blocks == 240
    e_d, m_d = _dev(eeg), _dev(emg)
    spec = torch.empty((210, 1, 100, 128), dtype=torch.complex64, device="cuda")
    plan.spectra(e_d, spec[..., :64], m_d, spec[..., 64:])
    torch.cuda.synchronize()
    got = spec[:, 0].cpu().numpy()
    ref = _oracle_spectra(np.concatenate([eeg, emg], axis=1), starts, 2048, 1, 1, 100)
    assert _rel_err(got, ref) < SPEC_TOL
    c_ref = oc.msc_from_spectra(ref[:, :, :64], ref[:, :, 64:])[0]
    c_got = oc.msc_from_spectra(got[:, :, :64].astype(np.complex128), got[:, :, 64:].astype(np.complex128))[0]
    assert np.max(np.abs(c_ref - c_got)) < 1e-5
    again = torch.empty_like(spec)
    plan.spectra(e_d, again[..., :64], m_d, again[..., 64:])
    g = torch.cuda.CUDAGraph()
    graphed = torch.empty_like(spec)
    with torch.cuda.graph(g):
        plan.spectra(e_d, graphed[..., :64], m_d, graphed[..., 64:])
    graphed.zero_()
    g.replay()
    g.replay()
    torch.cuda.synchronize()
    assert torch.equal(again.view(torch.float32), spec.view(torch.float32))
    assert torch.equal(graphed.view(torch.float32), spec.view(torch.float32))


def test_public_api_takes_the_tensor_core_path(cuda_device, monkeypatch):
    """welch_magnitude_squared_coherence on 64 x 64 channels: same coherence (1e-4 vs the fp64 oracle) whichever K1
    runs, the hann / band-limited call goes through the tensor cores, another window through the FFT kernel."""
    from multimodal_biosignal_analysis_b200 import kernels as K, signal_features as sf, synthetic as syn
    eeg, emg = syn.make_epochs(4, 8192, 64, 64, seed=9)
    starts = syn.epoch_segment_starts(4, 8192, 2048, 1024)
    taken = []
    orig = K.welch_spectra_pair
    monkeypatch.setattr(K, "welch_spectra_pair", lambda *a, **k: taken.append(orig(*a, **k)) or taken[-1])
    tc = sf.welch_magnitude_squared_coherence(eeg, emg, 2048.0, nperseg=2048, freq_band=(1, 100), segment_starts=starts)
    monkeypatch.setenv("CMC_WELCH_FFT", "1")
    fft = sf.welch_magnitude_squared_coherence(eeg, emg, 2048.0, nperseg=2048, freq_band=(1, 100), segment_starts=starts)
    monkeypatch.delenv("CMC_WELCH_FFT")
    ham = sf.welch_magnitude_squared_coherence(eeg, emg, 2048.0, nperseg=2048, window="hamming", freq_band=(1, 100),
                                               segment_starts=starts)
    assert taken == ["tensor-core", "fft", "fft"]
    ref = _oracle_spectra(np.concatenate([eeg, emg], axis=1), starts, 2048, 1, 1, 100)
    c_ref = oc.msc_from_spectra(ref[:, :, :64], ref[:, :, 64:])[0]
    assert np.max(np.abs(tc.coherence - c_ref)) < 1e-4
    assert np.max(np.abs(fft.coherence - c_ref)) < 1e-4
    assert np.max(np.abs(tc.coherence - fft.coherence)) < 2e-5
    assert ham.coherence.shape == tc.coherence.shape


def test_unsupported_requests_are_refused(cuda_device):
    from multimodal_biosignal_analysis_b200 import _lib, kernels as K
    assert not K.WelchHannPlan.supports(2048, 1, 200)          # band wider than 102 bins
    assert not K.WelchHannPlan.supports(1000, 1, 50)           # N not a multiple of 128
    assert K.hann_plan_for(np.arange(0, 4096, 1024), 2048, 1, 200) is None
    with pytest.raises(_lib.CmcError, match="outside the tensor-core kernel"):
        K.WelchHannPlan(np.arange(0, 4096, 1024), 2048, 1, 200)
    plan = K.WelchHannPlan(np.array([0, 1024], dtype=np.int64), 2048, 1, 100)
    x = torch.zeros((2048, 64), device="cuda")               # too short for the second segment
    out = torch.empty((2, 100, 64), dtype=torch.complex64, device="cuda")
    with pytest.raises(_lib.CmcError, match="outside the recording"):
        plan.spectra(x, out)


def test_cta_pair_variant_is_bit_identical(cuda_device, monkeypatch):
    """CMC_DT_PAIR=1 runs the same units as tcgen05 cta_group::2 pairs (M = 256, each CTA stages half of W): same
    accumulation order per output, so the spectra must be bit-identical - including an odd number of units (one
    empty unit in the last pair) and partly empty channel groups."""
    from multimodal_biosignal_analysis_b200 import kernels as K
    rng = np.random.default_rng(3)
    e = rng.standard_normal((9000, 20)).astype(np.float32)
    m = rng.standard_normal((9000, 64)).astype(np.float32)
    starts = np.array([0, 512, 1024, 1536, 4000, 6000, 6512], dtype=np.int64)      # 5 + 2 + 3 = 10 half blocks ...
    plan = K.WelchHannPlan(starts, 1024, 1, 60)
    starts_odd = starts[:-1]                                                          # ... and 5 + 2 + 2 = 9
    plan_odd = K.WelchHannPlan(starts_odd, 1024, 1, 60)
    assert plan.n_half_blocks == 10 and plan_odd.n_half_blocks == 9
    for pl in (plan, plan_odd):
        monkeypatch.delenv("CMC_DT_PAIR", raising=False)
        monkeypatch.setenv("CMC_DT_UNFOLD", "1")              # the unfolded single-CTA kernel: same arithmetic as the pair
        single = _run(K, pl, e, m, 60)
        monkeypatch.delenv("CMC_DT_UNFOLD")
        monkeypatch.setenv("CMC_DT_PAIR", "1")
        paired = _run(K, pl, e, m, 60)
        monkeypatch.delenv("CMC_DT_PAIR")
        assert not torch.isnan(torch.view_as_real(paired)).any()
        assert torch.equal(single.view(torch.float32), paired.view(torch.float32))


@pytest.mark.parametrize("seed", range(12))
def test_fuzz_shapes_bands_chains(cuda_device, seed):
    """Random segment length, band, channel counts, segment tables (chains of random length, isolated and duplicated
    segments, unsorted) and detrend modes against the fp64 oracle."""
    from multimodal_biosignal_analysis_b200 import kernels as K
    rng = np.random.default_rng(1000 + seed)
    N = int(rng.choice([256, 512, 1024, 2048, 4096, 1536, 384]))
    half = N // 2
    width = int(rng.integers(1, 103))
    b0_max = half - 104
    lo = int(rng.integers(0, max(1, min(b0_max, 300)) + 1))
    hi = min(lo + width - 1, half - 2)
    if hi + 1 - max(lo - 1, 0) > 103:
        hi = max(lo - 1, 0) + 102
    if max(lo - 1, 0) + 104 > half:
        lo, hi = 1, min(40, half - 106)
    assert K.WelchHannPlan.supports(N, lo, hi), (N, lo, hi)
    n1 = int(rng.integers(1, 131))
    n2 = int(rng.integers(0, 131))
    ld1, ld2 = (n1 + 3) // 4 * 4, (max(n2, 1) + 3) // 4 * 4          # TMA rows need 16-byte pitches
    starts = []
    pos = 0
    for _ in range(int(rng.integers(1, 6))):                          # chains
        pos += int(rng.integers(0, 3 * N))
        for k in range(int(rng.integers(1, 9))):
            starts.append(pos + k * half)
        pos = starts[-1] + N
    if rng.random() < 0.5:
        starts.append(starts[0])                                      # a duplicated segment
    starts = np.array(starts, dtype=np.int64)
    rng.shuffle(starts)
    n = int(starts.max()) + N + int(rng.integers(0, 50))
    scale = float(10.0 ** rng.uniform(-5, 3))                         # volts ... ADC counts
    x1 = (scale * (rng.standard_normal((n, ld1)) + rng.uniform(-3, 3, ld1))).astype(np.float32)
    x2 = (scale * (rng.standard_normal((n, ld2)) + rng.uniform(-3, 3, ld2))).astype(np.float32)
    detrend = int(rng.integers(0, 3))
    plan = K.WelchHannPlan(starts, N, lo, hi)
    F = hi - lo + 1
    x1d, x2d = _dev(x1)[:, :n1], _dev(x2)[:, :n2]
    spec = torch.full((len(starts), 1, F, n1 + n2 + 3), float("nan"), dtype=torch.complex64, device="cuda")
    if n2:
        plan.spectra(x1d, spec[..., :n1], x2d, spec[..., n1:n1 + n2], detrend=detrend)
    else:
        plan.spectra(x1d, spec[..., :n1], detrend=detrend)
    torch.cuda.synchronize()
    got = spec[:, 0, :, :n1 + n2].cpu().numpy()
    assert torch.isnan(spec[..., n1 + n2:].real).all()                                # nothing written past the channels
    ref = _oracle_spectra(np.concatenate([x1[:, :n1], x2[:, :n2]], axis=1), starts, N, detrend, lo, hi)
    assert not np.isnan(got.view(np.float32)).any()
    if detrend != 1 and lo == 0:
        # the DC bin of a non-detrended segment carries N/2 x mean: compare it relative to itself
        dc = np.abs(got[:, 0] - ref[:, 0]) / (np.abs(ref[:, 0]) + 1e-30)
        if detrend == 0:
            assert np.max(dc) < 1e-5
        got, ref = got[:, 1:], ref[:, 1:]
    if got.shape[1]:
        assert _rel_err(got, ref) < SPEC_TOL


def test_cluster_multicast_variant_is_bit_identical(cuda_device, monkeypatch):
    """CMC_DT_MC=1: clusters of two CTAs share every table k-block by TMA multicast (each CTA keeps its own units and
    MMAs).  Same arithmetic per output, so the spectra must be bit-identical to the default folded kernel - including an
    odd number of units (the rank-1 CTA of the last pair only keeps the table ring turning)."""
    from multimodal_biosignal_analysis_b200 import kernels as K
    rng = np.random.default_rng(5)
    e = rng.standard_normal((9000, 20)).astype(np.float32)
    m = rng.standard_normal((9000, 64)).astype(np.float32)
    starts = np.array([0, 512, 1024, 1536, 4000, 6000, 6512], dtype=np.int64)      # 10 half blocks
    for st in (starts, starts[:-1]):                                                  # ... and 9
        plan = K.WelchHannPlan(st, 1024, 1, 60)
        monkeypatch.delenv("CMC_DT_MC", raising=False)
        single = _run(K, plan, e, m, 60)
        monkeypatch.setenv("CMC_DT_MC", "1")
        shared = _run(K, plan, e, m, 60)
        monkeypatch.delenv("CMC_DT_MC")
        assert not torch.isnan(shared.real).any()
        assert torch.equal(single.view(torch.float32), shared.view(torch.float32))
