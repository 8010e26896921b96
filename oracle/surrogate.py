"""Surrogate-null oracle on cached spectra.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  PARITY UNPINNED: the reference
has no statistical surrogates (``src/pipeline/data_surrogation.py:19-199`` only
corrupts test data; its CMC significance threshold is the analytic
Beta(K-2, K-2) quantile, ``src/pipeline/signal_features.py:470-481``).  The two
constructions below are OUR definition of the north star's kernel (3); they are
the only ones for which a surrogate costs one cross-spectral pass over the
cached spectra:

  shift  Y_s[l] = Y[(l + shift_s * group) mod L]      (segment / window index
         rotation of one modality; ``group`` = tapers per window so whole
         windows move and taper indices stay aligned)
  phase  Y_s[l, f, :] = Y[l, f, :] * P[a(s, l, f)]    (one random phase per
         surrogate, segment and frequency, shared by all EMG channels, which
         preserves every auto-spectrum and the EMG inter-channel structure);
         a = Philox4x32-10(key = seed, counter = (s, l, f, 0))[0] >> 20 indexes
         a 4096-entry table P[a] = exp(2 pi i a / 4096) rounded to TF32.

  C_s = |sum_l conj(X) Y_s|^2 / (S_xx S_yy)   with the OBSERVED auto-spectra
  exceed[f,i,j] = #{s : C_s >= C_obs},  p = (1 + exceed) / (1 + n_surr)
  max_stat[s]   = max_{f,i,j} C_s
"""
from __future__ import annotations

import numpy as np

PHASE_BITS = 12
N_PHASES = 1 << PHASE_BITS

_M0 = np.uint64(0xD2511F53)
_M1 = np.uint64(0xCD9E8D57)
_W0 = np.uint32(0x9E3779B9)
_W1 = np.uint32(0xBB67AE85)
_MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Philox4x32-10 (Salmon et al. 2011), vectorised over equal-shaped uint32 counters."""
    c0 = np.asarray(c0, dtype=np.uint32).copy()
    c1 = np.asarray(c1, dtype=np.uint32).copy()
    c2 = np.asarray(c2, dtype=np.uint32).copy()
    c3 = np.asarray(c3, dtype=np.uint32).copy()
    k0 = np.uint32(k0)
    k1 = np.uint32(k1)
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0 = _M0 * c0.astype(np.uint64)
            p1 = _M1 * c2.astype(np.uint64)
            hi0 = (p0 >> np.uint64(32)).astype(np.uint32)
            lo0 = (p0 & _MASK).astype(np.uint32)
            hi1 = (p1 >> np.uint64(32)).astype(np.uint32)
            lo1 = (p1 & _MASK).astype(np.uint32)
            c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
            k0 = np.uint32((int(k0) + int(_W0)) & 0xFFFFFFFF)
            k1 = np.uint32((int(k1) + int(_W1)) & 0xFFFFFFFF)
    return c0, c1, c2, c3


def tf32_round(x: np.ndarray) -> np.ndarray:
    """Round-to-nearest-even to the 10-bit TF32 mantissa (finite float32 input)."""
    u = np.asarray(x, dtype=np.float32).view(np.uint32).astype(np.uint64)
    u = (u + np.uint64(0xFFF) + ((u >> np.uint64(13)) & np.uint64(1))) & np.uint64(0xFFFFE000)
    return u.astype(np.uint32).view(np.float32)


def phase_table() -> np.ndarray:
    """complex64 table P[a]; real and imaginary parts are TF32-representable."""
    ang = 2.0 * np.pi * np.arange(N_PHASES) / N_PHASES
    return (tf32_round(np.cos(ang).astype(np.float32))
            + 1j * tf32_round(np.sin(ang).astype(np.float32))).astype(np.complex64)


def phase_indices(seed: int, s: np.ndarray, L: int, F: int) -> np.ndarray:
    """int32 (len(s), L, F) table indices for surrogates ``s`` (global indices)."""
    s = np.asarray(s, dtype=np.uint32)
    S, Lg, Fg = np.meshgrid(s, np.arange(L, dtype=np.uint32), np.arange(F, dtype=np.uint32),
                            indexing="ij")
    r0, _, _, _ = philox4x32_10(S, Lg, Fg, np.zeros_like(S), seed & 0xFFFFFFFF,
                                (seed >> 32) & 0xFFFFFFFF)
    return (r0 >> np.uint32(32 - PHASE_BITS)).astype(np.int32)


def whiten(X: np.ndarray):
    """X (L, F, C) -> X / sqrt(sum_l |X|^2) (zero-power channels stay zero)."""
    p = np.sum(np.abs(X) ** 2, axis=0)
    scale = np.where(p > 0, 1.0 / np.sqrt(np.where(p > 0, p, 1.0)), 0.0)
    return X * scale[None], p


def surrogate_coherence(Xw, Yw, mode: str, s_index, shifts=None, group: int = 1, seed: int = 0):
    """fp64 coherence of the listed surrogates: (len(s_index), F, Ne, Nm)."""
    L, F, _ = Xw.shape
    out = []
    table = phase_table().astype(np.complex128)
    for s in np.asarray(s_index):
        if mode == "shift":
            Ys = np.roll(Yw, -int(shifts[s]) * group, axis=0)
        elif mode == "phase":
            a = phase_indices(seed, np.array([s]), L, F)[0]
            Ys = Yw * table[a][:, :, None]
        else:
            raise ValueError(mode)
        sxy = np.einsum("lfi,lfj->fij", np.conj(Xw), Ys)
        out.append(np.minimum(np.abs(sxy) ** 2, 1.0))
    return np.stack(out)


def null_statistics(coh_s: np.ndarray, coh_obs: np.ndarray, tol: float = 0.0):
    """exceedance counts (uint32 (F,Ne,Nm)) and per-surrogate max from a stack of
    surrogate coherences; ``tol`` shifts the comparison to build tolerance bands."""
    exceed = np.sum(coh_s >= (coh_obs[None] + tol), axis=0).astype(np.uint32)
    max_stat = coh_s.reshape(coh_s.shape[0], -1).max(axis=1)
    return exceed, max_stat
