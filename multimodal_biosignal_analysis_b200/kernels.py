"""Typed wrappers over the C ABI: torch CUDA tensors in, torch CUDA tensors out.

This is the only module that touches ctypes pointers.  Every function enqueues on the
current torch CUDA stream and never synchronises.
"""
from __future__ import annotations

import torch

from . import _lib

DETREND_NONE, DETREND_CONSTANT, DETREND_POST_TAPER = 0, 1, 2
SURR_SHIFT, SURR_PHASE = 0, 1
FIX_SHIFT = 30
FIX_SCALE = float(1 << FIX_SHIFT)


def _need_cuda(t: torch.Tensor, name: str, dtype=None):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise TypeError(f"{name} must be a CUDA tensor (this package has no CPU path)")
    if dtype is not None and t.dtype != dtype:
        raise TypeError(f"{name} must have dtype {dtype}, got {t.dtype}")


def check_segments(seg_starts_host, N: int, n_samples: int) -> None:
    """Host-side bounds check of a segment table (numpy / list)."""
    import numpy as np
    s = np.asarray(seg_starts_host)
    if s.size and (int(s.min()) < 0 or int(s.max()) + N > n_samples):
        raise ValueError("segment outside the recording")


def fft_segments(x: torch.Tensor, seg_starts: torch.Tensor, windows: torch.Tensor, detrend: int = 0,
                 bin_lo: int = 0, bin_hi: int | None = None, out: torch.Tensor | None = None,
                 ch_offset: int = 0) -> torch.Tensor:
    """Spectra of every (segment, window row, channel): complex64 (n_seg, K, F, C).

    x (n_samples, n_ch) float32 with unit channel stride; seg_starts int64 (n_seg,);
    windows float32 (K, N).  With ``out`` given (complex64 (n_seg, K, F, C_total)) the
    channels are written at columns [ch_offset, ch_offset + n_ch).
    """
    _need_cuda(x, "x", torch.float32)
    _need_cuda(seg_starts, "seg_starts", torch.int64)
    _need_cuda(windows, "windows", torch.float32)
    if x.dim() != 2 or x.stride(1) != 1:
        raise ValueError("x must be (n_samples, n_ch) with contiguous channels")
    windows = windows.contiguous()
    seg_starts = seg_starts.contiguous()
    n_samples, n_ch = x.shape
    K, N = windows.shape
    n_seg = seg_starts.numel()
    if bin_hi is None:
        bin_hi = N // 2
    F = bin_hi - bin_lo + 1
    # bounds of seg_starts are the caller's contract (checked on the host by the public API):
    # reading them back here would force a device synchronisation on every call
    if out is None:
        out = torch.empty((n_seg, K, F, n_ch), dtype=torch.complex64, device=x.device)
        ch_offset = 0
    else:
        _need_cuda(out, "out", torch.complex64)
        if out.shape[:3] != (n_seg, K, F) or not out.is_contiguous() or ch_offset + n_ch > out.shape[3]:
            raise ValueError("out has the wrong shape")
    lib = _lib.load()
    rc = lib.cmc_fft_segments(x.data_ptr(), n_samples, n_ch, x.stride(0), seg_starts.data_ptr(), n_seg,
                              windows.data_ptr(), K, N, detrend, bin_lo, bin_hi,
                              out.data_ptr() + 8 * ch_offset, out.shape[3], _lib.current_stream())
    _lib.check(rc, "cmc_fft_segments")
    return out


def fft_segments_pair(x1: torch.Tensor, x2: torch.Tensor, seg_starts: torch.Tensor, windows: torch.Tensor,
                      detrend: int, bin_lo: int, bin_hi: int, out1: torch.Tensor, out2: torch.Tensor) -> None:
    """Spectra of two recordings of equal length (EEG, EMG) in one K1 launch: ``out1`` / ``out2`` are complex64
    (n_seg, K, F, C_i) tensors or channel-range views of one (n_seg, K, F, C_total) array (same row pitch)."""
    for t, name in ((x1, "x1"), (x2, "x2")):
        _need_cuda(t, name, torch.float32)
        if t.dim() != 2 or t.stride(1) != 1:
            raise ValueError(f"{name} must be (n_samples, n_ch) with contiguous channels")
    if x1.shape[0] != x2.shape[0]:
        raise ValueError("both recordings must have the same number of samples")
    _need_cuda(seg_starts, "seg_starts", torch.int64)
    _need_cuda(windows, "windows", torch.float32)
    windows, seg_starts = windows.contiguous(), seg_starts.contiguous()
    K, N = windows.shape
    n_seg, F = seg_starts.numel(), bin_hi - bin_lo + 1
    for o, x, name in ((out1, x1, "out1"), (out2, x2, "out2")):
        _need_cuda(o, name, torch.complex64)
        if o.shape != (n_seg, K, F, x.shape[1]) or o.stride(3) != 1 or o.stride(1) != F * o.stride(2) or \
                o.stride(0) != K * F * o.stride(2):
            raise ValueError(f"{name} must be (n_seg, K, F, n_ch) with unit channel stride and dense leading axes")
    if out1.stride(2) != out2.stride(2):
        raise ValueError("out1 and out2 must share the row pitch")
    rc = _lib.load().cmc_fft_segments_pair(x1.data_ptr(), x1.shape[1], x1.stride(0), out1.data_ptr(),
                                           x2.data_ptr(), x2.shape[1], x2.stride(0), out2.data_ptr(),
                                           x1.shape[0], seg_starts.data_ptr(), n_seg, windows.data_ptr(), K, N, detrend,
                                           bin_lo, bin_hi, out1.stride(2), _lib.current_stream())
    _lib.check(rc, "cmc_fft_segments_pair")


class WelchHannPlan:
    """Half-block plan of the tensor-core Welch kernel (``cmc_welch_hann_*``): periodic hann window, one window row,
    band of at most 102 bins.  Built from the HOST segment table; ``spectra`` enqueues on the current stream."""

    def __init__(self, seg_starts_host, N: int, bin_lo: int, bin_hi: int):
        import ctypes as C
        import numpy as np
        starts = np.ascontiguousarray(np.asarray(seg_starts_host, dtype=np.int64))
        handle = C.c_void_p()
        lib = _lib.load()
        rc = lib.cmc_welch_hann_plan_create(starts.ctypes.data, int(starts.size), int(N), int(bin_lo), int(bin_hi),
                                            C.byref(handle))
        _lib.check(rc, "cmc_welch_hann_plan_create")
        self._handle, self._lib = handle, lib
        self.N, self.bin_lo, self.bin_hi, self.n_seg = int(N), int(bin_lo), int(bin_hi), int(starts.size)
        self.device = torch.cuda.current_device()
        nh = C.c_int()
        lib.cmc_welch_hann_plan_info(handle, C.byref(nh), None, None)
        self.n_half_blocks = int(nh.value)

    @staticmethod
    def supports(N: int, bin_lo: int, bin_hi: int) -> bool:
        b0 = max(bin_lo - 1, 0)
        return (N % 128 == 0 and 256 <= N <= 16384 and bin_hi + 1 - b0 <= 103 and b0 + 104 <= N // 2
                and 0 <= bin_lo <= bin_hi)

    def __del__(self):
        h = getattr(self, "_handle", None)
        if h is not None and h.value:
            try:
                self._lib.cmc_welch_hann_plan_destroy(h)
            except Exception:  # noqa: BLE001 - interpreter shutdown
                pass
            self._handle = None

    def spectra(self, x1: torch.Tensor, out1: torch.Tensor, x2: torch.Tensor | None = None,
                out2: torch.Tensor | None = None, detrend: int = DETREND_CONSTANT) -> None:
        """Spectra of the plan's segments for one recording or for two of equal length (EEG, EMG): ``out*`` are
        complex64 (n_seg, 1, F, C_i) or (n_seg, F, C_i) tensors / channel-range views that share the row pitch."""
        F = self.bin_hi - self.bin_lo + 1
        pairs = [(x1, out1, "1")] + ([(x2, out2, "2")] if x2 is not None else [])
        for x, o, name in pairs:
            _need_cuda(x, "x" + name, torch.float32)
            _need_cuda(o, "out" + name, torch.complex64)
            if x.dim() != 2 or x.stride(1) != 1:
                raise ValueError(f"x{name} must be (n_samples, n_ch) with contiguous channels")
            if o.dim() == 4:
                if o.shape[1] != 1:
                    raise ValueError("the tensor-core Welch kernel takes one window row")
                o = o[:, 0]
            if o.shape != (self.n_seg, F, x.shape[1]) or o.stride(2) != 1 or o.stride(0) != F * o.stride(1):
                raise ValueError(f"out{name} must be (n_seg, F, n_ch) with unit channel stride and dense leading axes")
        if x2 is not None and (x1.shape[0] != x2.shape[0] or out1.stride(-2) != out2.stride(-2)):
            raise ValueError("both recordings need the same length and both outputs the same row pitch")
        rc = self._lib.cmc_welch_hann_spectra(self._handle, x1.data_ptr(), x1.shape[1], x1.stride(0), out1.data_ptr(),
                                              _lib.ptr(x2), 0 if x2 is None else x2.shape[1],
                                              0 if x2 is None else x2.stride(0), _lib.ptr(out2), x1.shape[0],
                                              int(detrend), out1.stride(-2), _lib.current_stream())
        _lib.check(rc, "cmc_welch_hann_spectra")


_HANN_PLANS: dict = {}
_HANN_PLANS_MAX = 16


def hann_plan_for(seg_starts_host, N: int, bin_lo: int, bin_hi: int) -> "WelchHannPlan | None":
    """Cached :class:`WelchHannPlan` of a host segment table, or None when the tensor-core kernel does not take the
    request (band too wide, N not a multiple of 128, ``CMC_WELCH_FFT=1``)."""
    import os
    import numpy as np
    if os.environ.get("CMC_WELCH_FFT") or not WelchHannPlan.supports(N, bin_lo, bin_hi):
        return None
    starts = np.ascontiguousarray(np.asarray(seg_starts_host, dtype=np.int64))
    if starts.size == 0:
        return None
    key = (torch.cuda.current_device(), int(N), int(bin_lo), int(bin_hi), starts.size, hash(starts.tobytes()))
    plan = _HANN_PLANS.get(key)
    if plan is None:
        if torch.cuda.is_current_stream_capturing():
            return None                                   # plan creation allocates and synchronises
        if len(_HANN_PLANS) >= _HANN_PLANS_MAX:
            _HANN_PLANS.pop(next(iter(_HANN_PLANS)))
        plan = _HANN_PLANS[key] = WelchHannPlan(starts, N, bin_lo, bin_hi)
    return plan


def welch_spectra_pair(x1: torch.Tensor, x2: torch.Tensor, seg_starts_host, seg_starts: torch.Tensor,
                       windows: torch.Tensor, window_is_hann: bool, detrend: int, bin_lo: int, bin_hi: int,
                       out1: torch.Tensor, out2: torch.Tensor) -> str:
    """Spectra of two recordings (EEG, EMG) for the Welch entry points.  A single periodic-hann window row over a
    narrow band goes through the tensor-core half-block kernel (``cmc_welch_hann_spectra``), everything else through
    the FFT kernel (``cmc_fft_segments_pair``).  Returns the name of the path taken."""
    N = int(windows.shape[-1])
    use_tc = (window_is_hann and windows.shape[0] == 1 and x1.shape[1] + x2.shape[1] >= 64
              and x1.stride(0) % 4 == 0 and x2.stride(0) % 4 == 0 and x1.data_ptr() % 16 == 0 and x2.data_ptr() % 16 == 0)
    plan = hann_plan_for(seg_starts_host, N, bin_lo, bin_hi) if use_tc else None
    if plan is not None:
        plan.spectra(x1, out1, x2, out2, detrend=detrend)
        return "tensor-core"
    fft_segments_pair(x1, x2, seg_starts, windows, detrend, bin_lo, bin_hi, out1, out2)
    return "fft"


def psd_from_spectra(spec: torch.Tensor, base_scale: float, one_sided: bool, bin_lo: int, N: int,
                     log_scale: bool) -> torch.Tensor:
    """(W, K, F, C) complex64 spectra -> (W, F, C) float32 power spectra (mean over axis 1)."""
    _need_cuda(spec, "spec", torch.complex64)
    if spec.dim() != 4 or not spec.is_contiguous():
        raise ValueError("spec must be contiguous (W, K, F, C)")
    W, K, F, C = spec.shape
    out = torch.empty((W, F, C), dtype=torch.float32, device=spec.device)
    rc = _lib.load().cmc_psd_from_spectra(spec.data_ptr(), W, K, F, C, C, float(base_scale), int(one_sided),
                                          int(bin_lo), int(N), int(log_scale), out.data_ptr(), C,
                                          _lib.current_stream())
    _lib.check(rc, "cmc_psd_from_spectra")
    return out


def _spectra_dims(X, Y):
    _need_cuda(X, "X", torch.complex64)
    _need_cuda(Y, "Y", torch.complex64)
    if X.dim() != 4 or Y.dim() != 4 or X.shape[:3] != Y.shape[:3]:
        raise ValueError("X, Y must be (W, K, F, channels) with equal leading dims")
    if X.stride(3) != 1 or Y.stride(3) != 1:
        raise ValueError("channel stride must be 1")
    W, K, F, Ne = X.shape
    Nm = Y.shape[3]
    for t in (X, Y):
        if t.stride(0) != K * F * t.stride(2) or t.stride(1) != F * t.stride(2):
            raise ValueError("spectra must be dense in (W, K, F)")
    return W, K, F, Ne, Nm, X.stride(2), Y.stride(2)


def msc_windows(X: torch.Tensor, Y: torch.Tensor, window_mask: torch.Tensor | None = None,
                jackknife: bool = False, t_crit: float = 0.0, it_threshold: float | None = None):
    """Per-window MSC. Returns (coh, ci_lo, ci_hi, significant); absent outputs are None.
    Outputs of masked-out windows are zero."""
    W, K, F, Ne, Nm, ldx, ldy = _spectra_dims(X, Y)
    dev = X.device
    shape = (W, F, Ne, Nm)
    alloc = torch.zeros if window_mask is not None else torch.empty
    coh = alloc(shape, dtype=torch.float32, device=dev)
    lo = alloc(shape, dtype=torch.float32, device=dev) if jackknife else None
    hi = alloc(shape, dtype=torch.float32, device=dev) if jackknife else None
    sig = alloc(shape, dtype=torch.uint8, device=dev) if it_threshold is not None else None
    if window_mask is not None:
        _need_cuda(window_mask, "window_mask", torch.uint8)
        if window_mask.shape != (W,):
            raise ValueError("window_mask must have shape (W,)")
    lib = _lib.load()
    rc = lib.cmc_msc_windows(X.data_ptr(), Y.data_ptr(), W, K, F, Ne, Nm, ldx, ldy, _lib.ptr(window_mask),
                             int(jackknife), float(t_crit), -1.0 if it_threshold is None else float(it_threshold),
                             coh.data_ptr(), _lib.ptr(lo), _lib.ptr(hi), _lib.ptr(sig), _lib.current_stream())
    _lib.check(rc, "cmc_msc_windows")
    return coh, lo, hi, sig


def msc_windows_maxemg(X: torch.Tensor, Y: torch.Tensor, window_mask: torch.Tensor | None = None,
                       jackknife: bool = False, t_crit: float = 0.0, it_threshold: float | None = None,
                       zero_nonsignificant: bool = False, return_argmax: bool = False):
    """Fused per-window MSC + EMG-argmax: returns (coh, lo, hi, argmax) of shape (W, F, Ne)."""
    W, K, F, Ne, Nm, ldx, ldy = _spectra_dims(X, Y)
    dev = X.device
    shape = (W, F, Ne)
    alloc = torch.zeros if window_mask is not None else torch.empty
    coh = alloc(shape, dtype=torch.float32, device=dev)
    lo = alloc(shape, dtype=torch.float32, device=dev) if jackknife else None
    hi = alloc(shape, dtype=torch.float32, device=dev) if jackknife else None
    arg = alloc(shape, dtype=torch.int32, device=dev) if return_argmax else None
    if zero_nonsignificant and it_threshold is None:
        raise ValueError("zero_nonsignificant needs it_threshold")
    lib = _lib.load()
    rc = lib.cmc_msc_windows_maxemg(X.data_ptr(), Y.data_ptr(), W, K, F, Ne, Nm, ldx, ldy,
                                    _lib.ptr(window_mask), int(jackknife), float(t_crit),
                                    -1.0 if it_threshold is None else float(it_threshold),
                                    int(zero_nonsignificant), coh.data_ptr(), _lib.ptr(lo), _lib.ptr(hi),
                                    _lib.ptr(arg), _lib.current_stream())
    _lib.check(rc, "cmc_msc_windows_maxemg")
    return coh, lo, hi, arg


# ----------------------------------------------------------------------------- K2 / K3
def _pooled_dims(X, Y):
    _need_cuda(X, "X", torch.complex64)
    _need_cuda(Y, "Y", torch.complex64)
    if X.dim() != 3 or Y.dim() != 3 or X.shape[:2] != Y.shape[:2]:
        raise ValueError("X, Y must be (L, F, channels)")
    if X.stride(2) != 1 or Y.stride(2) != 1 or X.stride(0) != X.shape[1] * X.stride(1) or \
            Y.stride(0) != Y.shape[1] * Y.stride(1):
        raise ValueError("spectra must be dense in (L, F) with unit channel stride")
    L, F, Ne = X.shape
    return L, F, Ne, Y.shape[2], X.stride(1), Y.stride(1)


class PooledCsd:
    """Result of the tensor-core CSD pass; keeps (or can rebuild) the operands of the surrogate null.  Until the
    operand planes exist the object holds on to the spectra it was computed from (they are what the planes are
    built from on the first surrogate call)."""

    def __init__(self, coh, sxx, syy, sxy, ws, dims, pending=None):
        self.coh, self.sxx, self.syy, self.sxy, self.ws, self.dims = coh, sxx, syy, sxy, ws, dims
        # (X, Y, ldx, ldy) while the operand planes of ws are not filled yet
        self._pending = pending
        # ((seed, s_begin, s_end, f_begin, f_end), workspace) of the last phase null whose generated operands (phase
        # panel + cross-product rows) were kept for the histogram passes
        self.phase_ops = None
        self._xy = None             # the spectra the coherence was computed from (null_mean)

    def null_mean(self) -> torch.Tensor:
        """E[C_s] of the phase-randomised null per (f, i, j): sum_l |Xw_l|^2 |Yw_l|^2 for whitened spectra (the
        cross terms of |sum_l Z_l e^{i phi_l}|^2 vanish in expectation).  float32 (F, Ne, Nm)."""
        if self._xy is None:
            raise RuntimeError("the spectra of this result are not available")
        X, Y = self._xy
        px = X.real.square() + X.imag.square()
        py = Y.real.square() + Y.imag.square()
        wx = torch.where(self.sxx > 0, 1.0 / self.sxx, torch.zeros_like(self.sxx))
        wy = torch.where(self.syy > 0, 1.0 / self.syy, torch.zeros_like(self.syy))
        return torch.einsum("lfi,lfj->fij", px * wx[None], py * wy[None]).contiguous()

    def ensure_operands(self) -> None:
        """Fill the TF32 operand planes the surrogate kernels read (no-op when already present)."""
        if self._pending is None:
            return
        L, F, Ne, Nm = self.dims
        lib = _lib.load()
        ws_bytes = int(lib.cmc_csd_workspace_bytes(L, F, Ne, Nm))
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=self.coh.device)
        X, Y, ldx, ldy = self._pending
        rc = lib.cmc_csd_operands(X.data_ptr(), Y.data_ptr(), L, F, Ne, Nm, ldx, ldy, ws.data_ptr(), ws_bytes,
                                  _lib.current_stream())
        _lib.check(rc, "cmc_csd_operands")
        self.ws, self._pending = ws, None


def csd_msc(X: torch.Tensor, Y: torch.Tensor, want_sxy: bool = False, keep_operands: bool = False) -> PooledCsd:
    """Pooled coherence over the leading axis on the tensor cores: coh (F, Ne, Nm).

    By default only the coherence pass runs (the spectra are read once, straight from their (L, F, C) layout);
    the operand planes of the surrogate nulls are built on the first :func:`surrogate_null` call, or right away
    with ``keep_operands=True`` (pack + GEMM path)."""
    L, F, Ne, Nm, ldx, ldy = _pooled_dims(X, Y)
    dev = X.device
    lib = _lib.load()
    direct = (not keep_operands and ldx % 2 == 0 and ldy % 2 == 0 and X.data_ptr() % 16 == 0
              and Y.data_ptr() % 16 == 0)
    ws_bytes = int(lib.cmc_csd_workspace_bytes_min(F, Ne, Nm) if direct else lib.cmc_csd_workspace_bytes(L, F, Ne, Nm))
    if ws_bytes < 0:
        _lib.check(ws_bytes, "cmc_csd_workspace_bytes")
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    coh = torch.empty((F, Ne, Nm), dtype=torch.float32, device=dev)
    sxx = torch.empty((F, Ne), dtype=torch.float32, device=dev)
    syy = torch.empty((F, Nm), dtype=torch.float32, device=dev)
    sxy = torch.empty((F, Ne, Nm), dtype=torch.complex64, device=dev) if want_sxy else None
    fn = lib.cmc_csd_coherence if direct else lib.cmc_csd_msc
    rc = fn(X.data_ptr(), Y.data_ptr(), L, F, Ne, Nm, ldx, ldy, coh.data_ptr(), sxx.data_ptr(), syy.data_ptr(),
            _lib.ptr(sxy), ws.data_ptr(), ws_bytes, _lib.current_stream())
    _lib.check(rc, "cmc_csd_coherence" if direct else "cmc_csd_msc")
    out = PooledCsd(coh, sxx, syy, sxy, ws, (L, F, Ne, Nm), pending=(X, Y, ldx, ldy) if direct else None)
    out._xy = (X, Y)
    return out


def surrogate_null(csd: PooledCsd, mode: int, s_begin: int, s_end: int, shifts: torch.Tensor | None = None,
                   group: int = 1, seed: int = 0, exceed: torch.Tensor | None = None,
                   f_range: tuple[int, int] | None = None, keep_phase_operands: bool = False):
    """Surrogates [s_begin, s_end) against csd.coh: returns (exceed uint32 (F,Ne,Nm) accumulated,
    max_stat float32 (s_end - s_begin,)).  ``f_range = (f_begin, f_end)`` restricts the null to those
    frequency bins: exceed is only updated there and max_stat is the maximum over the range."""
    L, F, Ne, Nm = csd.dims
    dev = csd.coh.device
    n = s_end - s_begin
    lib = _lib.load()
    csd.ensure_operands()
    if exceed is None:
        exceed = torch.zeros((F, Ne, Nm), dtype=torch.int32, device=dev)
    max_stat = torch.empty(n, dtype=torch.float32, device=dev)
    if mode == SURR_SHIFT:
        _need_cuda(shifts, "shifts", torch.int32)
        if shifts.numel() != n:
            raise ValueError("one shift per surrogate in [s_begin, s_end)")
    ws2_bytes = int(lib.cmc_surrogate_workspace_bytes(L, F, Ne, Nm, mode, n))
    if ws2_bytes < 0:
        _lib.check(ws2_bytes, "cmc_surrogate_workspace_bytes")
    ws2 = torch.empty(max(ws2_bytes, 16), dtype=torch.uint8, device=dev)
    f_begin, f_end = (0, F) if f_range is None else (int(f_range[0]), int(f_range[1]))
    rc = lib.cmc_surrogate_null_range(csd.ws.data_ptr(), L, F, Ne, Nm, mode, group, _lib.ptr(shifts), seed, s_begin,
                                      s_end, f_begin, f_end, csd.coh.data_ptr(), exceed.data_ptr(),
                                      max_stat.data_ptr(), ws2.data_ptr(), ws2_bytes, _lib.current_stream())
    _lib.check(rc, "cmc_surrogate_null_range")
    if mode == SURR_PHASE:
        csd.phase_ops = ((int(seed), int(s_begin), int(s_end), f_begin, f_end), ws2) if keep_phase_operands else None
    return exceed, max_stat


def surrogate_null_hist(csd: PooledCsd, s_begin: int, s_end: int, seed: int = 0, n_bins: int = 128,
                        bin_lo: torch.Tensor | None = None, bin_scale: torch.Tensor | None = None,
                        hist: torch.Tensor | None = None, f_range: tuple[int, int] | None = None,
                        keep_operands: bool = False, below: torch.Tensor | None = None):
    """Per-pair histograms of the phase surrogates [s_begin, s_end): returns (hist int32 (F, Ne, Nm, n_bins), below
    int32 (F, Ne, Nm)), accumulated into ``hist`` / ``below`` when given.  Bin = floor((C_s - bin_lo) * bin_scale)
    with per-pair float32 (F, Ne, Nm) arrays (defaults 0 and n_bins: uniform bins over [0, 1]); values under the
    window are counted in ``below``, values over it are dropped."""
    L, F, Ne, Nm = csd.dims
    dev = csd.coh.device
    lib = _lib.load()
    csd.ensure_operands()
    if hist is None:
        hist = torch.zeros((F, Ne, Nm, n_bins), dtype=torch.int32, device=dev)
    elif hist.shape != (F, Ne, Nm, n_bins) or not hist.is_contiguous():
        raise ValueError("hist must be contiguous (F, Ne, Nm, n_bins)")
    if below is None:
        below = torch.zeros((F, Ne, Nm), dtype=torch.int32, device=dev)
    elif below.shape != (F, Ne, Nm) or not below.is_contiguous() or below.dtype != torch.int32:
        raise ValueError("below must be contiguous int32 (F, Ne, Nm)")
    for t, name in ((bin_lo, "bin_lo"), (bin_scale, "bin_scale")):
        if t is not None:
            _need_cuda(t, name, torch.float32)
            if t.shape != (F, Ne, Nm) or not t.is_contiguous():
                raise ValueError(f"{name} must be contiguous (F, Ne, Nm)")
    ws2_bytes = int(lib.cmc_surrogate_workspace_bytes(L, F, Ne, Nm, SURR_PHASE, s_end - s_begin))
    if ws2_bytes < 0:
        _lib.check(ws2_bytes, "cmc_surrogate_workspace_bytes")
    f_begin, f_end = (0, F) if f_range is None else (int(f_range[0]), int(f_range[1]))
    # the phase panel and the cross-product rows of an earlier pass over the same surrogates are reused (and kept
    # for the next zoom pass): they depend on (seed, surrogate range, frequency range) only
    key = (int(seed), int(s_begin), int(s_end), f_begin, f_end)
    reuse = csd.phase_ops is not None and csd.phase_ops[0] == key and csd.phase_ops[1].numel() >= ws2_bytes
    ws2 = csd.phase_ops[1] if reuse else torch.empty(max(ws2_bytes, 16), dtype=torch.uint8, device=dev)
    rc = lib.cmc_surrogate_null_hist(csd.ws.data_ptr(), L, F, Ne, Nm, SURR_PHASE, seed, s_begin, s_end, f_begin, f_end,
                                     n_bins, _lib.ptr(bin_lo), _lib.ptr(bin_scale), hist.data_ptr(), below.data_ptr(),
                                     ws2.data_ptr(), ws2.numel(), int(reuse), _lib.current_stream())
    _lib.check(rc, "cmc_surrogate_null_hist")
    csd.phase_ops = (key, ws2) if keep_operands else None
    return hist, below


def hist_select(hist: torch.Tensor, k: int, below: torch.Tensor) -> torch.Tensor:
    """Bin of the value of 0-based rank ``k`` in every histogram ``hist[..., :]`` whose counts start at ``below``
    (int32, same leading shape; updated in place to the count below the returned bin).  Returns int32 bins; -1 =
    the rank lies under the window, n_bins = over it."""
    _need_cuda(hist, "hist", torch.int32)
    _need_cuda(below, "below", torch.int32)
    if not hist.is_contiguous() or not below.is_contiguous() or below.shape != hist.shape[:-1]:
        raise ValueError("hist must be contiguous (..., n_bins) and below contiguous (...)")
    out = torch.empty_like(below)
    rc = _lib.load().cmc_hist_select(hist.data_ptr(), below.numel(), hist.shape[-1], int(k), below.data_ptr(),
                                     out.data_ptr(), _lib.current_stream())
    _lib.check(rc, "cmc_hist_select")
    return out


# ----------------------------------------------------------------------------- K4
def _cbpa_args(X, indptr, indices):
    _need_cuda(X, "X", torch.float64)
    _need_cuda(indptr, "indptr", torch.int32)
    _need_cuda(indices, "indices", torch.int32)
    if X.dim() != 2 or not X.is_contiguous():
        raise ValueError("X must be contiguous (n_subj, n_tests)")
    n_subj, n_tests = X.shape
    if indptr.numel() != n_tests + 1:
        raise ValueError("indptr must have n_tests + 1 entries")
    return n_subj, n_tests


def cbpa_workspace(X: torch.Tensor) -> torch.Tensor:
    """Scratch for the CBPA calls on X; pass it to both calls to share the re-tiled copy of X."""
    n_subj, n_tests = X.shape
    return torch.empty(int(_lib.load().cmc_cbpa_workspace_bytes(n_subj, n_tests)), dtype=torch.uint8, device=X.device)


def cbpa_observed(X: torch.Tensor, thr: float, tail: int, indptr: torch.Tensor, indices: torch.Tensor,
                  ws: torch.Tensor | None = None):
    """Observed clustering: (t_obs f64 (n_tests,), labels int32, mass_fixed int64 (n_clusters,),
    n_clusters).  Synchronises once to read the cluster count.  Leaves the tiled copy of X in ``ws``."""
    n_subj, n_tests = _cbpa_args(X, indptr, indices)
    dev = X.device
    lib = _lib.load()
    if ws is None:
        ws = cbpa_workspace(X)
    ws_bytes = ws.numel()
    t_obs = torch.empty(n_tests, dtype=torch.float64, device=dev)
    labels = torch.empty(n_tests, dtype=torch.int32, device=dev)
    mass_fixed = torch.zeros(n_tests, dtype=torch.int64, device=dev)
    mass_f64 = torch.zeros(n_tests, dtype=torch.float64, device=dev)
    n_clusters = torch.zeros(1, dtype=torch.int32, device=dev)
    rc = lib.cmc_cbpa_observed(X.data_ptr(), n_subj, n_tests, float(thr), int(tail), indptr.data_ptr(),
                               indices.data_ptr(), t_obs.data_ptr(), labels.data_ptr(), mass_fixed.data_ptr(),
                               mass_f64.data_ptr(), n_clusters.data_ptr(), ws.data_ptr(), ws_bytes,
                               _lib.current_stream())
    _lib.check(rc, "cmc_cbpa_observed")
    n = int(n_clusters.item())
    return t_obs, labels, mass_fixed[:n], n


def cbpa_permute(X: torch.Tensor, signs: torch.Tensor, p_begin: int, p_end: int, thr: float, tail: int,
                 indptr: torch.Tensor, indices: torch.Tensor, ws: torch.Tensor | None = None,
                 tiled: bool = False) -> torch.Tensor:
    """Max-cluster statistic (int64 fixed point) of permutations [p_begin, p_end) of the sign table.
    ``tiled=True``: ``ws`` already holds the tiled copy of X (left by :func:`cbpa_observed` or an earlier call on
    the same stream), the re-tile pass is skipped."""
    n_subj, n_tests = _cbpa_args(X, indptr, indices)
    _need_cuda(signs, "signs", torch.int8)
    if signs.dim() != 2 or signs.shape[1] != n_subj or not signs.is_contiguous():
        raise ValueError("signs must be contiguous int8 (n_perm, n_subj)")
    if not (0 <= p_begin <= p_end <= signs.shape[0]):
        raise ValueError("permutation range outside the sign table")
    h0 = torch.empty(p_end - p_begin, dtype=torch.int64, device=X.device)
    if p_end == p_begin:                        # a rank whose shard of a short (exact) sign table is empty
        return h0
    lib = _lib.load()
    if tiled and ws is None:
        raise ValueError("tiled=True needs the workspace that holds the tiled copy")
    if ws is None:
        ws = cbpa_workspace(X)
    ws_bytes = ws.numel()
    rc = lib.cmc_cbpa_permute(None if tiled else X.data_ptr(), n_subj, n_tests, signs.data_ptr(), p_begin, p_end, float(thr),
                              int(tail), indptr.data_ptr(), indices.data_ptr(), h0.data_ptr(), ws.data_ptr(),
                              ws_bytes, _lib.current_stream())
    _lib.check(rc, "cmc_cbpa_permute")
    return h0
