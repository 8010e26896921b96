"""Seeded synthetic EEG / HD-EMG / CBPA inputs.

Shared by the tests and ``bench.py`` so that the CUDA path, the oracle and the
CPU baseline all see the same float32 values.  Pure numpy/scipy, no GPU.

Signal model: a common beta/gamma source s(t) (two AR(2) resonators at 20 Hz and
40 Hz driven by white noise); ``eeg[:, i] = a_i s + pink_i`` and
``emg[:, j] = b_j s(t - tau) + white_j`` with tau = 12 ms; 8 EEG and 16 EMG
channels are coupled (gains ~ U(0.1, 0.6)), the rest carry noise only.  Pink
noise follows the reference's own recipe (white rFFT scaled by 1/sqrt(f),
``src/pipeline/data_surrogation.py:174-186``).
"""
from __future__ import annotations

import numpy as np
from scipy import signal

FS = 2048.0


def _resonator(freq_hz: float, r: float, fs: float):
    w = 2 * np.pi * freq_hz / fs
    return [1.0], [1.0, -2 * r * np.cos(w), r * r]


def _pink(rng: np.random.Generator, n: int, n_ch: int) -> np.ndarray:
    w = np.fft.rfft(rng.standard_normal((n, n_ch)), axis=0)
    f = np.fft.rfftfreq(n)
    f[0] = 1.0
    p = np.fft.irfft(w / np.sqrt(f)[:, None], n=n, axis=0)
    return p / p.std(axis=0, keepdims=True)


def make_recording(n_samples: int, n_eeg: int = 64, n_emg: int = 64, seed: int = 20260102,
                   fs: float = FS, n_coupled_eeg: int = 8, n_coupled_emg: int = 16,
                   dtype=np.float32):
    """Continuous recording: returns (eeg (n, n_eeg), emg (n, n_emg)) time-first."""
    rng = np.random.default_rng(seed)
    tau = int(round(0.012 * fs))
    drive = rng.standard_normal(n_samples + tau + 512)
    s = np.zeros_like(drive)
    for f0 in (20.0, 40.0):
        b, a = _resonator(f0, 0.985, fs)
        s += signal.lfilter(b, a, drive)
    s = s[512:] / s[512:].std()
    s_eeg = s[tau:tau + n_samples]
    s_emg = s[:n_samples]                      # delayed by tau relative to EEG
    a_gain = np.zeros(n_eeg)
    b_gain = np.zeros(n_emg)
    ce = rng.permutation(n_eeg)[:min(n_coupled_eeg, n_eeg)]
    cm = rng.permutation(n_emg)[:min(n_coupled_emg, n_emg)]
    a_gain[ce] = rng.uniform(0.1, 0.6, len(ce))
    b_gain[cm] = rng.uniform(0.1, 0.6, len(cm))
    eeg = s_eeg[:, None] * a_gain[None, :] + _pink(rng, n_samples, n_eeg)
    emg = s_emg[:, None] * b_gain[None, :] + rng.standard_normal((n_samples, n_emg))
    return np.ascontiguousarray(eeg, dtype=dtype), np.ascontiguousarray(emg, dtype=dtype)


def make_epochs(n_epochs: int = 30, epoch_samples: int = 8192, n_eeg: int = 64, n_emg: int = 64,
                seed: int = 20260102, dtype=np.float32):
    """BASELINE config 2: ``n_epochs`` task epochs of 4 s at 2048 Hz, laid out as one
    contiguous (n_epochs * epoch_samples, n_ch) array; epoch e owns rows
    [e * epoch_samples, (e + 1) * epoch_samples)."""
    return make_recording(n_epochs * epoch_samples, n_eeg, n_emg, seed=seed, dtype=dtype)


def epoch_segment_starts(n_epochs: int, epoch_samples: int, nperseg: int, hop: int) -> np.ndarray:
    """Welch segment starts that never straddle an epoch boundary (7 per 4-s epoch for
    nperseg 2048 / hop 1024)."""
    per = (epoch_samples - nperseg) // hop + 1
    base = np.arange(per, dtype=np.int64) * hop
    return (np.arange(n_epochs, dtype=np.int64)[:, None] * epoch_samples + base[None, :]).reshape(-1)


# 2-D sensor layout used for the synthetic CBPA adjacency: 64 points on a jittered
# polar grid (the real montage coordinates are an input of the product, not
# something it needs to reproduce - SURVEY.md 8c).
def sensor_positions(n_ch: int = 64, seed: int = 7) -> np.ndarray:
    rng = np.random.default_rng(seed)
    rings = [1, 6, 12, 18, 27]
    pts = []
    for ri, cnt in enumerate(rings):
        rad = ri / (len(rings) - 1)
        for k in range(cnt):
            ang = 2 * np.pi * (k + 0.5 * (ri % 2)) / cnt
            pts.append((rad * np.cos(ang), rad * np.sin(ang)))
    pts = np.asarray(pts[:n_ch]) + rng.normal(0, 0.01, (n_ch, 2))
    return pts


def make_cbpa_contrast(n_subj: int = 20, n_times: int = 100, n_ch: int = 64, seed: int = 20260104,
                       effect: float = 0.8):
    """BASELINE config 4: X (n_subj, n_times, n_ch) = N(0,1) + effect on a block of
    3 lattice bins x 10 neighbouring channels."""
    rng = np.random.default_rng(seed)
    X = rng.standard_normal((n_subj, n_times, n_ch))
    t0 = n_times // 3
    X[:, t0:t0 + 3, : min(10, n_ch)] += effect
    return X


def make_sign_table(n_perm: int, n_subj: int, seed: int = 42, tail: int = 0) -> np.ndarray:
    """Host-supplied sign-flip table int8 (n_perm, n_subj) in {-1,+1}; column 0 is
    forced to +1 for two-tailed tests (exploits symmetry like MNE does)."""
    rng = np.random.default_rng(seed)
    order = rng.integers(0, 2, (n_perm, n_subj), dtype=np.int8)
    if tail == 0:
        order[:, 0] = 1
    return (2 * order - 1).astype(np.int8)
