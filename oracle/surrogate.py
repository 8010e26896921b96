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
         a = Philox4x32-10(key = seed, counter = (s, l, f >> 2, 0))[f & 3] >> 20 indexes
         the 4096 unit-circle phases P[a] = exp(2 pi i a / 4096), evaluated in float64.

The definition is UNQUANTISED float64 arithmetic.  The CUDA path rounds its tensor-core
operands (shift: 3xTF32 split; phase: FP16 cross-products and FP16 phase table); that
rounding is kernel error, measured by the tests against this definition with the north
star's 1e-4 gate.  ``emulate_fp16`` below only exists to attribute that error.

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


def f16_round(x: np.ndarray) -> np.ndarray:
    """Round-to-nearest-even to IEEE binary16, returned as float64."""
    return np.asarray(x, dtype=np.float64).astype(np.float16).astype(np.float64)


Z_PRESCALE = 2.0 ** 14      # the kernel stores 2^14 Z in FP16 (|Z| <= 1 for whitened spectra)


def phase_table() -> np.ndarray:
    """complex128 P[a] = exp(2 pi i a / 4096): the definition."""
    ang = 2.0 * np.pi * np.arange(N_PHASES) / N_PHASES
    return np.cos(ang) + 1j * np.sin(ang)


def kernel_phase_table() -> np.ndarray:
    """The FP16-rounded table the CUDA kernel multiplies with (diagnostics; equals ``cmc_phase_table``)."""
    t = phase_table()
    return f16_round(t.real) + 1j * f16_round(t.imag)


def phase_indices(seed: int, s: np.ndarray, L: int, F: int) -> np.ndarray:
    """int32 (len(s), L, F) table indices for surrogates ``s`` (global indices)."""
    s = np.asarray(s, dtype=np.uint32)
    n_groups = (F + 3) // 4
    S, Lg, Fg = np.meshgrid(s, np.arange(L, dtype=np.uint32), np.arange(n_groups, dtype=np.uint32),
                            indexing="ij")
    words = philox4x32_10(S, Lg, Fg, np.zeros_like(S), seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    idx = np.stack(words, axis=-1).reshape(len(s), L, n_groups * 4)[:, :, :F]
    return (idx >> np.uint32(32 - PHASE_BITS)).astype(np.int32)


def whiten(X: np.ndarray):
    """X (L, F, C) -> X / sqrt(sum_l |X|^2) (zero-power channels stay zero)."""
    p = np.sum(np.abs(X) ** 2, axis=0)
    scale = np.where(p > 0, 1.0 / np.sqrt(np.where(p > 0, p, 1.0)), 0.0)
    return X * scale[None], p


def surrogate_coherence(Xw, Yw, mode: str, s_index, shifts=None, group: int = 1, seed: int = 0, table=None,
                        emulate_fp16: bool = False):
    """fp64 coherence of the listed surrogates: (len(s_index), F, Ne, Nm).

    ``emulate_fp16`` (phase mode, DIAGNOSTIC ONLY - not the definition): round the cross-products
    Z[l,f,i,j] = conj(Xw) Yw (prescaled by 2^14) and the phase table to FP16 like the single-term
    tensor-core operands the CUDA kernel uses for L > 85, to attribute its deviation from the
    definition (shorter averages run a hi/lo split that is exact to ~1e-6)."""
    L, F, _ = Xw.shape
    out = []
    if table is None:
        table = kernel_phase_table() if emulate_fp16 else phase_table()
    table = np.asarray(table).astype(np.complex128)
    Z = None
    if mode == "phase" and emulate_fp16:
        Z = np.conj(Xw)[:, :, :, None] * Yw[:, :, None, :]
        Z = (f16_round(Z.real * Z_PRESCALE) + 1j * f16_round(Z.imag * Z_PRESCALE)) / Z_PRESCALE
    for s in np.asarray(s_index):
        if mode == "shift":
            Ys = np.roll(Yw, -int(shifts[s]) * group, axis=0)
            sxy = np.einsum("lfi,lfj->fij", np.conj(Xw), Ys)
        elif mode == "phase":
            a = phase_indices(seed, np.array([s]), L, F)[0]
            if Z is not None:
                sxy = np.einsum("lfij,lf->fij", Z, table[a])
            else:
                sxy = np.einsum("lfi,lfj->fij", np.conj(Xw), Yw * table[a][:, :, None])
        else:
            raise ValueError(mode)
        out.append(np.minimum(np.abs(sxy) ** 2, 1.0))
    return np.stack(out)


def null_statistics(coh_s: np.ndarray, coh_obs: np.ndarray, tol: float = 0.0):
    """exceedance counts (uint32 (F,Ne,Nm)) and per-surrogate max from a stack of
    surrogate coherences; ``tol`` shifts the comparison to build tolerance bands."""
    exceed = np.sum(coh_s >= (coh_obs[None] + tol), axis=0).astype(np.uint32)
    max_stat = coh_s.reshape(coh_s.shape[0], -1).max(axis=1)
    return exceed, max_stat
