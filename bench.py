#!/usr/bin/env python
"""Benchmark of the cortico-muscular coherence hot path on B200.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # CPU reference arm (oracle port)

Headline metric (BASELINE.json): EEG x EMG coherence pair-spectra/s on config 2 - one subject-
condition = 64-ch EEG x 64-ch HD-EMG, 30 task epochs of 4 s at 2048 Hz, Welch segments of 2048
samples / hop 1024 (L = 210), 1-100 Hz band (F = 100) -> 4,096 pair-spectra per step.
One step = K1 (detrend + hann + rFFT of every segment of all 128 channels) + K2 (one kernel: spectra
staged as MN-major operands, TF32 split, auto-spectra, tcgen05 CSD -> MSC).  Every rank processes its own subject-condition per step (weak scaling, no
data-path collective).  ``value`` has the inputs resident in HBM; ``e2e`` goes through the public
Python API with pinned host buffers (H2D of the recording and D2H of the coherence inside the
timed region).  The surrogate-null and CBPA stages are timed separately and reported under
``stages`` (they shard a fixed total over the ranks: strong scaling).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FS = 2048.0
N_EPOCHS, EPOCH = 30, 8192
NPERSEG, HOP = 2048, 1024
BAND = (1.0, 100.0)
NE = NM = 64
N_ROTATE = 4                 # distinct resident recordings rotated per step (4 x 126 MB > 126 MB L2)
N_SURR = 1000                # config 3
N_PERM_TOTAL = 10000         # config 5 permutation count (sharded over ranks)
CBPA_SHAPE = (20, 100, 64)   # config 4


def _k1_traffic(kind="tc"):
    """dram__bytes_read.sum + dram__bytes_write.sum of one K1 launch from the committed `ncu --set full` export
    (profiled offline, NOT measured in this run; null when the file is absent).  kind: "tc" | "fft"."""
    path = os.path.join(ROOT, "profiles", "k1_dram_traffic.json")
    try:
        with open(path) as fh:
            d = json.load(fh)[kind]
        return {"bytes": float(d["dram_bytes_read"]) + float(d["dram_bytes_write"]),
                "source": f"profiled offline: profiles/k1_dram_traffic.json ({d.get('source', '')})"}
    except Exception:
        return {"bytes": None, "source": "no ncu export committed"}




def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            p = json.load(fh)
        return float(p["hbm_gbs"]), float(p.get("bf16_tflops", 1590.0)), "measured (MEASURED_PEAKS.json)"
    return 6650.0, 1590.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                 "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx = float(r[2])
            except (ValueError, IndexError):
                continue
            for name, col in (("hw_slowdown", 4), ("hw_thermal_slowdown", 5), ("sw_thermal_slowdown", 6),
                              ("sw_power_cap", 7)):
                if len(r) > col and r[col].lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------ CPU arm
def _cpu_fft_chunk(args):
    from oracle import coherence as oc
    x, starts, win, lo, hi = args
    return oc.segment_spectra(x, starts, win, 1, lo, hi)[:, 0]


def _cpu_csd_chunk(args):
    from oracle import coherence as oc
    X, Y = args
    return oc.msc_from_spectra(X, Y)[0]


def cpu_pooled_coherence(eeg, emg, starts, pool, n_workers):
    """Oracle (numpy fp64) Welch all-pairs coherence, parallel over segments (FFT) and frequency
    chunks (CSD) with ``n_workers`` processes."""
    from scipy import signal
    win = signal.get_window("hann", NPERSEG)[None]
    freqs = np.fft.rfftfreq(NPERSEG, 1 / FS)
    sel = np.flatnonzero((freqs >= BAND[0]) & (freqs <= BAND[1]))
    lo, hi = int(sel[0]), int(sel[-1])
    chunks = np.array_split(starts, min(n_workers, len(starts)))
    jobs = [(eeg, c, win, lo, hi) for c in chunks] + [(emg, c, win, lo, hi) for c in chunks]
    parts = pool.map(_cpu_fft_chunk, jobs) if pool else [_cpu_fft_chunk(j) for j in jobs]
    X = np.concatenate(parts[:len(chunks)])
    Y = np.concatenate(parts[len(chunks):])
    fch = np.array_split(np.arange(X.shape[1]), min(n_workers, X.shape[1]))
    jobs = [(X[:, f], Y[:, f]) for f in fch]
    cs = pool.map(_cpu_csd_chunk, jobs) if pool else [_cpu_csd_chunk(j) for j in jobs]
    return np.concatenate(cs)


def run_cpu_reference(steps, warmup, sample_epochs=N_EPOCHS):
    """Times the oracle port on the host cores; returns (pair-spectra/s, ms/step, cores, sample)."""
    import multiprocessing as mp
    from multimodal_biosignal_analysis_b200 import synthetic as syn
    cores = os.cpu_count() or 1
    eeg, emg = syn.make_epochs(sample_epochs, EPOCH, NE, NM, seed=20260102)
    eeg, emg = eeg.astype(np.float64), emg.astype(np.float64)
    starts = syn.epoch_segment_starts(sample_epochs, EPOCH, NPERSEG, HOP)
    ctx = mp.get_context("fork")
    pool = ctx.Pool(cores) if cores > 1 else None
    try:
        for _ in range(warmup):
            cpu_pooled_coherence(eeg, emg, starts, pool, cores)
        t0 = time.perf_counter()
        for _ in range(steps):
            cpu_pooled_coherence(eeg, emg, starts, pool, cores)
        dt = (time.perf_counter() - t0) / max(steps, 1)
    finally:
        if pool:
            pool.terminate()
    # cost is linear in the number of segments: scale a sub-sampled run to the full L = 210
    scale = N_EPOCHS / sample_epochs
    ms = dt * 1e3 * scale
    sample = (f"oracle port (numpy fp64 Welch all-pairs) on {sample_epochs}/{N_EPOCHS} epochs of config 2, "
              f"{cores} worker processes" + (", time scaled linearly to 30 epochs" if scale != 1 else ""))
    return NE * NM / (ms / 1e3), ms, cores, sample


_CPU_STAGE = {}


def _cpu_surr_chunk(args):
    from oracle import surrogate as osur
    mode, idx, shifts, seed = args
    Xw, Yw, coh_obs = _CPU_STAGE["Xw"], _CPU_STAGE["Yw"], _CPU_STAGE["coh"]
    cs = osur.surrogate_coherence(Xw, Yw, mode, idx, shifts=shifts, seed=seed)
    cnt, ms = osur.null_statistics(cs, coh_obs)
    return cnt, ms


def _cpu_cbpa_chunk(args):
    from oracle import cbpa as ocb
    from scipy import sparse
    signs, thr = args
    Xf, coo = _CPU_STAGE["Xf"], _CPU_STAGE["coo"]
    out = []
    for sg in signs:
        tp = ocb.ttest_1samp_no_p(Xf * sg[:, None].astype(np.float64))
        _, _, im = ocb.find_clusters(tp, thr, 0, coo)
        out.append(ocb.max_stat_fixed(im, 0))
    return out


def _cpu_mt_chunk(args):
    from oracle import coherence as oc
    starts, lo, hi = args
    eeg, emg, tapers = _CPU_STAGE["eeg"], _CPU_STAGE["emg"], _CPU_STAGE["tapers"]
    n = 0
    for s0 in starts:
        X = oc.segment_spectra(eeg, np.array([s0]), tapers, 0, lo, hi)[0]          # (K, F, Ne), no detrend
        Y = oc.segment_spectra(emg, np.array([s0]), tapers, 0, lo, hi)[0]
        m, l, h = oc.jackknife_from_spectra(X, Y, 0.05)
        n += int((m > 0.81).sum() >= 0)                                            # mask as in the GPU stage
    return n


def run_cpu_stage_baselines(n_surr_cpu=32, n_perm_cpu=256, n_win_cpu=32):
    """Oracle ports of the surrogate null and the CBPA on the host cores (bounded samples, forked workers share
    the inputs): returns {stage: cpu_baseline dict}."""
    import multiprocessing as mp
    from scipy import signal, sparse
    from scipy.stats import t as t_dist
    from oracle import cbpa as ocb, coherence as oc, surrogate as osur
    from multimodal_biosignal_analysis_b200 import synthetic as syn
    cores = os.cpu_count() or 1
    out = {}
    # spectra of config 2 (fp64), whitened once - the CPU analogue of the cached operands
    eeg, emg = syn.make_epochs(N_EPOCHS, EPOCH, NE, NM, seed=20260102)
    starts = syn.epoch_segment_starts(N_EPOCHS, EPOCH, NPERSEG, HOP)
    win = signal.get_window("hann", NPERSEG)[None]
    freqs = np.fft.rfftfreq(NPERSEG, 1 / FS)
    sel = np.flatnonzero((freqs >= BAND[0]) & (freqs <= BAND[1]))
    lo, hi = int(sel[0]), int(sel[-1])
    X = oc.segment_spectra(eeg.astype(np.float64), starts, win, 1, lo, hi)[:, 0]
    Y = oc.segment_spectra(emg.astype(np.float64), starts, win, 1, lo, hi)[:, 0]
    _CPU_STAGE["Xw"], _ = osur.whiten(X)
    _CPU_STAGE["Yw"], _ = osur.whiten(Y)
    _CPU_STAGE["coh"] = oc.msc_from_spectra(X, Y)[0]
    Xc = syn.make_cbpa_contrast(*CBPA_SHAPE)
    adj = ocb.combine_adjacency(CBPA_SHAPE[1], ocb.delaunay_adjacency(syn.sensor_positions(CBPA_SHAPE[2])))
    _CPU_STAGE["Xf"] = np.ascontiguousarray(Xc.reshape(CBPA_SHAPE[0], -1), dtype=np.float64)
    _CPU_STAGE["coo"] = sparse.coo_matrix(adj)
    thr = float(t_dist.ppf(0.975, CBPA_SHAPE[0] - 1))
    signs = syn.make_sign_table(n_perm_cpu, CBPA_SHAPE[0], seed=42)
    shifts = np.random.default_rng(3).integers(1, len(starts), n_surr_cpu).astype(np.int32)
    _CPU_STAGE["eeg"], _CPU_STAGE["emg"] = eeg.astype(np.float64), emg.astype(np.float64)
    _CPU_STAGE["tapers"] = oc.dpss_tapers(NPERSEG, 3, 0.9)[0]
    ctx = mp.get_context("fork")
    pool = ctx.Pool(cores) if cores > 1 else None
    mapper = pool.map if pool else (lambda f, jobs: [f(j) for j in jobs])
    try:
        chunks = [c for c in np.array_split(starts[:n_win_cpu], cores) if len(c)]
        jobs = [(c, lo, hi) for c in chunks]
        mapper(_cpu_mt_chunk, jobs[:1])
        t0 = time.perf_counter()
        mapper(_cpu_mt_chunk, jobs)
        dt = time.perf_counter() - t0
        out["multitaper_windows"] = {
            "value": min(n_win_cpu, len(starts)) * NE * NM / dt, "unit": "pair-spectra/s (one per window)",
            "cores": cores, "kind": "port",
            "sample": f"oracle/coherence.py (DPSS spectra + leave-one-taper-out jackknife CI, F = {hi - lo + 1} in-band "
                      f"bins): {min(n_win_cpu, len(starts))} of the 210 windows, {cores} worker processes"}
        for mode, key in (("phase", "surrogate_null_phase"), ("shift", "surrogate_null_shift")):
            chunks = [c for c in np.array_split(np.arange(n_surr_cpu), cores) if len(c)]
            jobs = [(mode, c, shifts if mode == "shift" else None, 7) for c in chunks]
            mapper(_cpu_surr_chunk, jobs[:1])                                  # warm the workers
            t0 = time.perf_counter()
            mapper(_cpu_surr_chunk, jobs)
            dt = time.perf_counter() - t0
            out[key] = {"value": n_surr_cpu / dt, "unit": "surrogates/s", "cores": cores, "kind": "port",
                        "sample": f"oracle/surrogate.py ({mode}): {n_surr_cpu} surrogates of config 2 (64x64xF=100, "
                                  f"L={len(starts)}) on cached whitened spectra, numpy complex128 einsum per surrogate, "
                                  f"{cores} worker processes"}
        chunks = [c for c in np.array_split(np.arange(n_perm_cpu), cores) if len(c)]
        jobs = [(signs[c], thr) for c in chunks]
        mapper(_cpu_cbpa_chunk, jobs[:1])
        t0 = time.perf_counter()
        mapper(_cpu_cbpa_chunk, jobs)
        dt = time.perf_counter() - t0
        out["cbpa"] = {"value": n_perm_cpu / dt, "unit": "permutations/s", "cores": cores, "kind": "port",
                       "sample": f"oracle/cbpa.py (MNE algorithm: numpy t-map + scipy connected_components): "
                                 f"{n_perm_cpu} permutations of the config 4 geometry, {cores} worker processes"}
    finally:
        if pool:
            pool.terminate()
        _CPU_STAGE.clear()
    return out


def main_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warmup = max(args.steps, 1), max(args.warmup, 1)
    # bounded sample: calibrate on one epoch, then pick the number of epochs per step so that the whole
    # run (warm-up + K steps) stays near two minutes of CPU wall time
    _, ms1, _, _ = run_cpu_reference(1, 1, sample_epochs=1)
    per_epoch_s = ms1 / 1e3 / N_EPOCHS
    sample_epochs = int(max(1, min(N_EPOCHS, 120.0 / ((steps + warmup) * per_epoch_s))))
    val, ms, cores, sample = run_cpu_reference(steps, warmup, sample_epochs=sample_epochs)
    line = {
        "impl": "reference", "metric": "pair_spectra_per_s", "value": val, "unit": "pair-spectra/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(),
        "cpu_baseline": {"value": val, "unit": "pair-spectra/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "pair-spectra/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def workload_config():
    return {"workload": "BASELINE config 2: 64x64 all-pairs Welch MSC, 30 epochs x 4 s @ 2048 Hz, nperseg 2048 "
                        "hop 1024 (L=210 segments), 1-100 Hz (F=100) = 4096 pair-spectra per subject-condition "
                        "per step per GPU",
            "l2_policy": f"inputs larger than L2: {N_ROTATE} resident recordings (126 MB each) rotated per step",
            "parallelism": "one subject-condition per rank per step, no data-path collective",
            "pipelining": "one K1 launch per step transforms the EEG and the EMG array; consecutive steps overlap on "
                          "two streams: K2 of step i (normal priority) runs beside K1 of step i + 1 (high priority); "
                          "the timed region replays CUDA graphs of 20 such steps"}


# ------------------------------------------------------------------------------------------ GPU arm
def main_gpu(args):
    import torch
    import torch.distributed as dist
    from scipy import signal
    from scipy.stats import t as t_dist
    from multimodal_biosignal_analysis_b200 import _lib, kernels as K, synthetic as syn
    from multimodal_biosignal_analysis_b200 import signal_features as sf
    from multimodal_biosignal_analysis_b200 import data_surrogation as dsur
    from multimodal_biosignal_analysis_b200 import cbpa as cb
    from multimodal_biosignal_analysis_b200 import dist as cdist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    # stdout carries the ONE JSON line: whatever libraries print on fd 1 (NCCL's version banner) goes to stderr
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    _lib.load()
    steps, warmup = max(args.steps, 1), max(args.warmup, 3)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms: float) -> float:
        if world == 1:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- resident synthetic inputs (different per rank and per rotation slot) ----
    starts_h = syn.epoch_segment_starts(N_EPOCHS, EPOCH, NPERSEG, HOP)
    starts = torch.from_numpy(starts_h).to(dev)
    win = torch.from_numpy(signal.get_window("hann", NPERSEG).astype(np.float32)[None]).to(dev)
    freqs = np.fft.rfftfreq(NPERSEG, 1 / FS)
    sel = np.flatnonzero((freqs >= BAND[0]) & (freqs <= BAND[1]))
    lo, hi = int(sel[0]), int(sel[-1])
    F = hi - lo + 1
    L = len(starts_h)
    host_sets, dev_sets = [], []
    for r in range(N_ROTATE):
        eeg, emg = syn.make_epochs(N_EPOCHS, EPOCH, NE, NM, seed=20260102 + 97 * rank + r)
        host_sets.append((eeg, emg))
        dev_sets.append((torch.from_numpy(eeg).to(dev), torch.from_numpy(emg).to(dev)))
    specs = [torch.empty((L, 1, F, NE + NM), dtype=torch.complex64, device=dev) for _ in range(N_ROTATE)]

    # K1: hann / 50 % overlap / narrow band = the tensor-core half-block kernel (cmc_welch_hann_spectra); --k1 fft
    # times the FFT kernel (cmc_fft_segments_pair) instead.  The plan is built once, outside every timed region.
    k1_plan = None if args.k1 == "fft" else K.hann_plan_for(starts_h, NPERSEG, lo, hi)
    K1_TRAFFIC_PROFILED = _k1_traffic("tc" if k1_plan is not None else "fft")
    k1_name = ("dft_hann_tc_kernel (K1t, tcgen05 BF16x3 half-block DFT, ONE launch for the EEG and the EMG array)"
               if k1_plan is not None else
               "fft_segments_tma_pipe_kernel<1024> (K1, ONE launch for the EEG and the EMG array)")

    def k1(eeg_d, emg_d, spec):
        if k1_plan is not None:
            k1_plan.spectra(eeg_d, spec[..., :NE], emg_d, spec[..., NE:], detrend=K.DETREND_CONSTANT)
        else:
            K.fft_segments_pair(eeg_d, emg_d, starts, win, K.DETREND_CONSTANT, lo, hi, spec[..., :NE], spec[..., NE:])

    def step(i):
        eeg_d, emg_d = dev_sets[i % N_ROTATE]
        spec = specs[i % N_ROTATE]
        k1(eeg_d, emg_d, spec)
        return K.csd_msc(spec[:, 0, :, :NE], spec[:, 0, :, NE:])

    for i in range(warmup):
        res = step(i)
    torch.cuda.synchronize()
    # One CUDA-graph pair per rotation slot: gA = the two K1 launches, gB = the direct K2 kernel (TMA of the
    # spectra as MN-major operands, TF32 split + auto-spectra in shared memory, tcgen05 GEMM, coherence epilogue).
    # The steps are software-pipelined over two streams: K2 of recording i (stream B) runs while K1 of recording
    # i + 1 (stream A) already occupies the SMs it leaves idle (100 tiles on 148 SMs); K1 claims its tiles from a
    # device-wide counter, so CTAs that start late simply take fewer.  Every slot has its own spectra buffer.
    graphs = []
    side = torch.cuda.Stream()
    for r in range(N_ROTATE):
        eeg_d, emg_d = dev_sets[r]
        spec = specs[r]
        gA, gB = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
        n0 = _lib.launch_count()
        with torch.cuda.graph(gA):
            # ONE K1 launch transforms both modalities (tiles of the EEG and of the EMG array share the persistent CTAs
            # and the claim counter: one prologue and one tail instead of two)
            k1(eeg_d, emg_d, spec)
        with torch.cuda.graph(gB):
            res_r = K.csd_msc(spec[:, 0, :, :NE], spec[:, 0, :, NE:])
        launches_per_step = _lib.launch_count() - n0
        graphs.append((gA, gB, res_r))
    for i in range(max(warmup, N_ROTATE)):
        graphs[i % N_ROTATE][0].replay()
        graphs[i % N_ROTATE][1].replay()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    # (Without the stream priorities below, one graph of 20 pipelined steps was SLOWER than per-step graphs, 56.8 against
    # 50.2 us per step, and so were two steps per graph and stream, 53.7 us: K2 then took its 100 SMs ahead of the next
    # K1.  Per-step graphs cost the host ~37 us per 50 us step - too close to being host bound on a slower machine.)
    TIMED = 4
    # K1 runs on a HIGH-priority stream, K2 on a normal one: when K1 of step i ends, K2 of step i and K1 of step i + 1
    # become ready together, and K1 - 148 CTAs with a static two-round schedule - must get the SMs first; K2's 100 CTAs
    # then start on the 56 SMs K1 leaves idle in its second round.  With --graph-steps G (default 20) the timed region
    # replays CUDA graphs of G steps captured on those two streams (kernel nodes keep the priority of the stream they
    # were captured on), so the host enqueues one graph per G steps; K mod G steps run through the per-step graphs.
    sA, sB = torch.cuda.Stream(priority=-1), torch.cuda.Stream()
    G = max(0, args.graph_steps)
    n_big, rem = (steps // G, steps % G) if G else (0, steps)
    big = None
    if n_big:
        big = torch.cuda.CUDAGraph()
        with torch.cuda.graph(big, stream=sA):
            ev_b = []
            for j in range(G):
                if j >= N_ROTATE:
                    sA.wait_event(ev_b[j - N_ROTATE])              # the slot's spectra buffer has been consumed
                k1(*dev_sets[j % N_ROTATE], specs[j % N_ROTATE])
                e = torch.cuda.Event()
                e.record(sA)
                sB.wait_event(e)
                with torch.cuda.stream(sB):
                    K.csd_msc(specs[j % N_ROTATE][:, 0, :, :NE], specs[j % N_ROTATE][:, 0, :, NE:])
                    e2 = torch.cuda.Event()
                    e2.record(sB)
                    ev_b.append(e2)
            sA.wait_stream(sB)
        big.replay()                                               # untimed first replay
        torch.cuda.synchronize()
    a_done = [torch.cuda.Event(enable_timing=(i % TIMED == 0)) for i in range(rem)]
    b_done = [torch.cuda.Event(enable_timing=(i % TIMED == 0)) for i in range(rem)]
    a_start = {i: torch.cuda.Event(enable_timing=True) for i in range(0, rem, TIMED)}
    b_start = {i: torch.cuda.Event(enable_timing=True) for i in range(0, rem, TIMED)}
    t_begin = torch.cuda.Event(enable_timing=True)
    t_end = torch.cuda.Event(enable_timing=True)
    barrier()
    cur = torch.cuda.current_stream()
    sA.wait_stream(cur)
    sB.wait_stream(cur)
    with torch.cuda.stream(sA):
        t_begin.record()
    host_t0 = time.perf_counter()
    with torch.cuda.stream(sA):
        for _ in range(n_big):
            big.replay()
    if n_big:
        sB.wait_stream(sA)
    for i in range(rem):
        gA, gB, res = graphs[i % N_ROTATE]
        with torch.cuda.stream(sA):
            if i >= N_ROTATE:
                sA.wait_event(b_done[i - N_ROTATE])     # the slot's spectra buffer has been consumed
            if i in a_start:
                a_start[i].record()
            gA.replay()
            a_done[i].record()
        with torch.cuda.stream(sB):
            sB.wait_event(a_done[i])
            if i in b_start:
                b_start[i].record()
            gB.replay()
            b_done[i].record()
    host_enqueue_ms = (time.perf_counter() - host_t0) * 1e3 / steps      # must stay below ms_per_step
    sB.wait_stream(sA)
    with torch.cuda.stream(sB):
        t_end.record()                                  # stream B finishes last (it waits for stream A)
    cur.wait_stream(sA)
    cur.wait_stream(sB)
    barrier()
    launches = launches_per_step * steps
    total_ms = max_over_ranks(t_begin.elapsed_time(t_end))
    clocks = sampler.stop() if rank == 0 else None
    ms_per_step = total_ms / steps
    value = NE * NM * world / (ms_per_step / 1e3)
    del a_start, b_start
    # per-kernel durations for the roofline: right after the timed region (same process, same clocks, inputs rotated
    # the same way) two CUDA graphs, one with 2 * N_ROTATE K1 launches and one with as many K2 launches, are replayed
    # on ONE stream with CUDA events around every replay: launches of one stream do not overlap, and inside a graph no
    # host launch gap sits between them, so elapsed / launches is the average launch duration of that kernel
    R = 2 * N_ROTATE
    g_k1, g_k2 = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
    with torch.cuda.graph(g_k1):
        for j in range(R):
            k1(*dev_sets[j % N_ROTATE], specs[j % N_ROTATE])
    with torch.cuda.graph(g_k2):
        for j in range(R):
            K.csd_msc(specs[j % N_ROTATE][:, 0, :, :NE], specs[j % N_ROTATE][:, 0, :, NE:])
    for g_ in (g_k1, g_k2):
        g_.replay()
    torch.cuda.synchronize()
    n_serial = max(4, min(steps, 128) // R)
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(n_serial)]
    for i in range(n_serial):
        ev[i][0].record()
        g_k1.replay()
        ev[i][1].record()
        g_k2.replay()
        ev[i][2].record()
    torch.cuda.synchronize()
    k1_ms = float(np.mean([e[0].elapsed_time(e[1]) for e in ev])) / R         # the K1 launch (both modalities)
    k2_ms = float(np.mean([e[1].elapsed_time(e[2]) for e in ev])) / R
    serial_ms_per_step = k1_ms + k2_ms
    n_samples = N_EPOCHS * EPOCH
    k1_bytes = n_samples * (NE + NM) * 4 + L * F * (NE + NM) * 8                # per launch (both modalities)
    hbm, bf16, peak_src = peaks()
    k1_gbs = k1_bytes / (k1_ms * 1e-3) / 1e9

    # ---- end to end through the public API: pinned host buffers in, numpy coherence out ----
    numa_node = cdist.bind_to_gpu_numa_node(local_rank) if world > 1 else None   # node-local pinned buffers per rank
    pinned = [(torch.from_numpy(a).pin_memory(), torch.from_numpy(b).pin_memory()) for a, b in host_sets]

    def e2e_step(i):
        eeg_h, emg_h = pinned[i % N_ROTATE]
        pc = sf.welch_magnitude_squared_coherence(eeg_h, emg_h, FS, nperseg=NPERSEG, freq_band=BAND,
                                                  segment_starts=starts_h)
        return pc.coherence                                                  # D2H, synchronises

    def e2e_sweep(n):
        # the sweep API: upload of recording i + 1, K1 + K2 of recording i, download of recording i - 1 overlap
        chk = 0.0
        for coh, _ in sf.welch_coherence_sweep((pinned[i % N_ROTATE] for i in range(n)), FS, nperseg=NPERSEG,
                                               freq_band=BAND, segment_starts=starts_h):
            chk += float(coh[0, 0, 0])                                       # the result is read on the host
        return chk

    for i in range(2):
        coh_e2e = e2e_step(i)
    e2e_sweep(4)
    barrier()
    t0 = time.perf_counter()
    e2e_sweep(steps)
    barrier()
    e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3) / steps
    n_single = min(steps, 50)
    t0 = time.perf_counter()
    for i in range(n_single):
        coh_e2e = e2e_step(i)
    single_ms = (time.perf_counter() - t0) * 1e3 / n_single
    e2e = {"value": NE * NM * world / (e2e_ms / 1e3), "unit": "pair-spectra/s", "ms_per_step": e2e_ms,
           "h2d_bytes_per_step": int(n_samples * (NE + NM) * 4), "d2h_bytes_per_step": int(F * NE * NM * 4),
           "h2d_gb_per_s_per_gpu": n_samples * (NE + NM) * 4 / (e2e_ms * 1e6),
           "bound": "host->device copy: the step is the upload of one recording over the GPU's PCIe link",
           "api": "signal_features.welch_coherence_sweep(recordings, ...): one recording per step, pinned host "
                  "tensors in, numpy coherence out; upload, K1 + K2 and download of consecutive recordings overlap",
           "single_call_ms": single_ms, "numa_node": numa_node,
           "single_call_api": "signal_features.welch_magnitude_squared_coherence(...).coherence, one blocking call "
                              "per recording"}

    stages = {}
    if not args.skip_stages:
        # ---- stage: surrogate null (config 3: 1,000 circular-shift surrogates on the cached spectra) ----
        shifts = np.random.default_rng(3).integers(1, L, N_SURR).astype(np.int32)
        shifts_d = torch.from_numpy(shifts).to(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n_rep = 3
        for _ in range(2):
            K.surrogate_null(res, K.SURR_SHIFT, 0, N_SURR, shifts=shifts_d)
        barrier()
        e0.record()
        for _ in range(n_rep):
            ex_s, ms_s = K.surrogate_null(res, K.SURR_SHIFT, 0, N_SURR, shifts=shifts_d)
        e1.record()
        barrier()
        surr_ms = max_over_ranks(e0.elapsed_time(e1)) / n_rep
        n_distinct = len(np.unique(shifts))
        flop = 3 * 2.0 * 128 * 64 * 2 * L * F * n_distinct                   # executed TF32 flop (three terms)
        stages["surrogate_null_shift"] = {
            "metric": "surrogates_per_s", "value": N_SURR / (surr_ms / 1e3), "unit": "surrogates/s",
            "ms": surr_ms, "scaling": "replicated",
            # what the kernel actually contracts: one CSD pass per DISTINCT shift (the p-value resolution of a shift
            # null is 1 / (L - 1) however many surrogates are drawn)
            "distinct_shifts": int(n_distinct), "csd_passes_per_s": n_distinct / (surr_ms / 1e3),
            "config": f"config 3: {N_SURR} circular-shift surrogates of one 64x64xF=100 subject-condition; distinct shifts are "
                      f"deduplicated on the device ({n_distinct} of L={L}), four of them share one 3xTF32 tcgen05 tile "
                      f"(N = 256), so the cost does not grow beyond {L - 1} CSD passes (10,000 surrogates take the same "
                      f"time); every rank runs it for its own subject-condition",
            "roofline": {"bound": "tensor", "achieved": flop / (surr_ms * 1e-3) / 1e12, "peak": bf16 / 2,
                         "unit": "TFLOP/s", "frac": flop / (surr_ms * 1e-3) / 1e12 / (bf16 / 2),
                         "note": "executed TF32 flop of the distinct-shift passes (M=128 x N=256 tiles, three error-compensated "
                                 "TF32 terms per product); TF32 peak taken as half the measured dense bf16 figure"},
        }

        # ---- stage: phase-randomised surrogates (config 3 count per rank-shard of config 5's 10,000) ----
        n_phase = 10000
        # one null shared by the ranks: every rank runs all surrogates on its slice of the frequency axis, so
        # operand generation, phases and contraction all shrink with the rank count (data_surrogation._plan)
        fb, fe = cdist.shard_range(F, rank, world)
        def phase_null():
            ex_p, ms_p = K.surrogate_null(res, K.SURR_PHASE, 0, n_phase, seed=7, f_range=(fb, fe))
            # one all-gather carries every rank's slice of the counts together with its per-surrogate maxima
            return cdist.all_gather_frequency_slices(ex_p, ms_p, (fb, fe))

        for _ in range(2):                                             # warm-up includes the collectives
            phase_null()
        barrier()
        e0.record()
        for _ in range(n_rep):
            ex_p, ms_all = phase_null()
        e1.record()
        barrier()
        ph_ms = max_over_ranks(e0.elapsed_time(e1)) / n_rep
        kpb = ((2 * L + 63) // 64) * 64
        ph_flop = 2.0 * n_phase * (2 * NE * NM) * kpb * (fe - fb)              # executed fp16 flop on this rank
        stages["surrogate_null_phase"] = {
            "metric": "surrogates_per_s", "value": n_phase / (ph_ms / 1e3), "unit": "surrogates/s", "ms": ph_ms,
            "scaling": "strong",
            "config": f"config 5 count: {n_phase} phase-randomised surrogates of one 64x64xF=100 subject-condition "
                      f"with the frequency axis sharded over {world} rank(s) (Philox phases + fp16 Z operands generated in "
                      f"the timed region; "
                      f"count slices and per-surrogate maxima exchanged with ONE all-gather)",
            "roofline": {"bound": "tensor", "achieved": ph_flop / (ph_ms * 1e-3) / 1e12, "peak": bf16,
                         "unit": "TFLOP/s", "frac": ph_flop / (ph_ms * 1e-3) / 1e12 / bf16,
                         "note": "executed fp16 flop (kind::f16, K padded to 64) vs the measured dense bf16 peak (same rate)"},
        }

        # ---- stage: config 3 with per-pair significance thresholds (null histograms, two zoom passes) ----
        import types
        _Pooled = types.SimpleNamespace(device_result=res, coherence=res.coh, freqs=freqs[lo:hi + 1], group=1)
        for _ in range(2):
            dsur.phase_randomised_surrogate_null(_Pooled, N_SURR, seed=3, thresholds=True)
        barrier()
        t0 = time.perf_counter()
        for _ in range(n_rep):
            null3 = dsur.phase_randomised_surrogate_null(_Pooled, N_SURR, seed=3, thresholds=True)
        barrier()
        thr_ms = max_over_ranks((time.perf_counter() - t0) * 1e3) / n_rep
        stages["surrogate_thresholds_cfg3"] = {
            "metric": "surrogates_per_s", "value": N_SURR / (thr_ms / 1e3), "unit": "surrogates/s", "ms": thr_ms,
            "scaling": "strong" if world > 1 else "replicated",
            "significant_pairs": int(null3["significant"].sum()),
            "config": f"config 3: {N_SURR} phase-randomised surrogates per pair of one 64x64xF=100 subject-condition -> "
                      "exceedance p-values, family-wise threshold AND per-pair (1 - alpha) thresholds from device-side "
                      "null histograms (128 bins, two zoom passes = three GEMM sweeps in total); wall clock of "
                      "data_surrogation.phase_randomised_surrogate_null(..., thresholds=True) incl. the download of "
                      "counts, maxima and thresholds"}

        # ---- stage: the reference's production estimator - per-window multitaper MSC with jackknife CI ----
        from multimodal_biosignal_analysis_b200.signal_features import _dpss
        tapers = torch.from_numpy(_dpss(NPERSEG, 3, 0.9).astype(np.float32)).to(dev)
        Kt = tapers.shape[0]
        t_crit = float(t_dist.ppf(0.975, Kt - 1))
        eeg_d, emg_d = dev_sets[0]

        def mt_step():
            S = torch.empty((L, Kt, F, NE + NM), dtype=torch.complex64, device=dev)
            K.fft_segments_pair(eeg_d, emg_d, starts, tapers, K.DETREND_NONE, lo, hi, S[..., :NE], S[..., NE:])
            return K.msc_windows(S[..., :NE], S[..., NE:], None, True, t_crit, 0.81)

        for _ in range(3):
            mt_step()                                     # warm the allocator: 1.1 GB of outputs per call
        barrier()
        e0.record()
        for _ in range(n_rep):
            mt_step()
        e1.record()
        barrier()
        mt_ms = max_over_ranks(e0.elapsed_time(e1)) / n_rep
        mt_bytes = n_samples * (NE + NM) * 4 + L * F * NE * NM * 13 + 2 * L * Kt * F * (NE + NM) * 8
        stages["multitaper_windows"] = {
            "metric": "window_pair_spectra_per_s", "value": L * NE * NM * world / (mt_ms / 1e3),
            "unit": "pair-spectra/s (one per window)", "ms": mt_ms, "scaling": "weak",
            "config": f"multitaper variant of config 2: {L} windows of {NPERSEG} samples, K={Kt} DPSS tapers, 64x64 pairs, "
                      f"F={F} in-band bins, jackknife CI + independence mask (signal_features.py:619-839 semantics); "
                      f"outputs (W,F,64,64) x (3 float32 + 1 mask) stay in HBM",
            "roofline": {"bound": "hbm", "achieved": mt_bytes / (mt_ms * 1e-3) / 1e9, "peak": hbm, "unit": "GB/s",
                         "frac": mt_bytes / (mt_ms * 1e-3) / 1e9 / hbm,
                         "note": "algorithmic bytes = recordings once + spectra write/read + 13 B per (window, bin, pair)"},
        }

        # ---- stage: CBPA permutations (config 4 geometry, config 5 count sharded over the ranks) ----
        from multimodal_biosignal_analysis_b200.cbpa import combine_adjacency, find_ch_adjacency_from_positions
        adj = combine_adjacency(CBPA_SHAPE[1], find_ch_adjacency_from_positions(syn.sensor_positions(CBPA_SHAPE[2])))
        adj.sort_indices()
        Xc = syn.make_cbpa_contrast(*CBPA_SHAPE)
        signs = syn.make_sign_table(N_PERM_TOTAL, CBPA_SHAPE[0], seed=42)
        thr = float(t_dist.ppf(0.975, CBPA_SHAPE[0] - 1))
        Xd = torch.from_numpy(np.ascontiguousarray(Xc.reshape(CBPA_SHAPE[0], -1))).to(dev)
        indptr = torch.from_numpy(adj.indptr.astype(np.int32)).to(dev)
        indices = torch.from_numpy(adj.indices.astype(np.int32)).to(dev)
        sd = torch.from_numpy(signs).to(dev)
        pb, pe = cdist.shard_range(N_PERM_TOTAL, rank, world)
        ws_c = K.cbpa_workspace(Xd)
        K.cbpa_observed(Xd, thr, 0, indptr, indices, ws=ws_c)          # observed clustering; leaves the tiled X in ws_c
        for _ in range(2):                                             # warm-up includes the collective
            cdist.all_gather_ranges(K.cbpa_permute(Xd, sd, pb, pe, thr, 0, indptr, indices, ws=ws_c, tiled=True),
                                    N_PERM_TOTAL)
        barrier()
        e0.record()
        for _ in range(n_rep):
            h0 = cdist.all_gather_ranges(K.cbpa_permute(Xd, sd, pb, pe, thr, 0, indptr, indices, ws=ws_c, tiled=True),
                                         N_PERM_TOTAL)
        e1.record()
        barrier()
        cbpa_ms = max_over_ranks(e0.elapsed_time(e1)) / n_rep
        # the t-map is the dominant pass and lives in the FP64 pipe: numpy's operation order costs 4 n_subj - 2
        # adds/multiplies + 4 divisions/square root per test, none of them fusable; X is L2 resident
        n_subj_c, n_tests_c = CBPA_SHAPE[0], CBPA_SHAPE[1] * CBPA_SHAPE[2]
        fp64_ops = n_tests_c * (4 * n_subj_c - 2 + 4)
        cbpa_tops = fp64_ops * (pe - pb) / (cbpa_ms * 1e-3) / 1e12
        n_sm = torch.cuda.get_device_properties(dev).multi_processor_count
        sm_clock_hz = torch.cuda.get_device_properties(dev).clock_rate * 1e3 if hasattr(
            torch.cuda.get_device_properties(dev), "clock_rate") else 1.965e9
        fp64_peak = n_sm * 64 * sm_clock_hz / 1e12       # FP64 instructions: 64 lanes / SM / clock at the device's boost clock
        stages["cbpa"] = {
            "metric": "cbpa_permutations_per_s", "value": N_PERM_TOTAL / (cbpa_ms / 1e3), "unit": "permutations/s",
            "ms": cbpa_ms, "scaling": "strong",
            "config": f"config 4 geometry (20 subj x 100 x 64 = 6400 tests, {adj.nnz} nnz adjacency), "
                      f"{N_PERM_TOTAL} sign-flip permutations sharded over {world} rank(s), H0 all-gathered",
            "roofline": {"bound": "fp64", "achieved": cbpa_tops, "peak": fp64_peak, "unit": "Tinstr/s",
                         "frac": cbpa_tops / fp64_peak,
                         "note": "algorithmic FP64 operations of the sign-flip t-map (4 n_subj + 2 per test, unfused as "
                                 "in numpy) vs the FP64 issue rate of 148 SMs x 64 lanes x 1965 MHz; X (1 MB) is L2 "
                                 "resident, DRAM traffic is negligible"},
        }

        # ---- stage: BASELINE config 5 as one pipeline through the public API (numpy / pinned host in, numpy out) ----
        # 20 subjects x 4 conditions = 80 subject-conditions dealt round-robin over the ranks (no collective while they
        # run); per unit upload -> K1 x 2 -> K2 -> operand planes -> 10,000-surrogate phase null -> download of coherence,
        # counts and maxima, uploads / downloads of neighbouring units overlapped; then one all-reduce of the (F, Ne)
        # EMG-max maps and the 10,000-permutation CBPA of the condition contrast (permutations sharded, H0 all-gathered).
        from multimodal_biosignal_analysis_b200 import sweep as csweep
        n_subj5, conds5 = 20, ("happy", "sad", "calm", "silence")
        units5 = {(f"S{s_:02d}", c_): pinned[(4 * s_ + k_) % N_ROTATE]
                  for s_ in range(n_subj5) for k_, c_ in enumerate(conds5)}

        def run_cfg5(units, n_surr=10000, n_perm=N_PERM_TOTAL):
            return csweep.cmc_surrogate_cbpa_sweep(units, FS, nperseg=NPERSEG, freq_band=BAND, segment_starts=starts_h,
                                                   n_surrogates=n_surr, mode="phase", seed=11, n_permutations=n_perm,
                                                   contrasts=[("happy", "silence")])

        small = {k: v for k, v in list(units5.items())[: 4 * max(world, 2)]}
        run_cfg5(small)                                                # warm-up: allocator, NCCL, CBPA tables
        barrier()
        t0 = time.perf_counter()
        out5 = run_cfg5(units5)
        barrier()
        cfg5_s = max_over_ranks((time.perf_counter() - t0) * 1e3) / 1e3
        n_units5 = len(units5)
        # the unit stage alone on THIS rank's share (device events around one more pass over 8 resident units)
        e0.record()
        for _ in range(8):
            K.surrogate_null(res, K.SURR_PHASE, 0, 10000, seed=3)
        e1.record()
        torch.cuda.synchronize()
        null_ms = e0.elapsed_time(e1) / 8
        stages["config5_sweep"] = {
            "metric": "subject_conditions_per_s", "value": n_units5 / cfg5_s, "unit": "subject-conditions/s",
            "seconds_total": cfg5_s, "ms_per_unit_wall": cfg5_s * 1e3 * world / n_units5, "scaling": "strong",
            "surrogates_per_s_e2e": n_units5 * 10000 / cfg5_s,
            "pair_spectra_per_s_e2e": n_units5 * NE * NM / cfg5_s,
            "null_kernel_ms_per_unit": null_ms,
            "h2d_bytes_per_unit": int(n_samples * (NE + NM) * 4),
            "d2h_bytes_per_unit": int(F * NE * NM * 8 + 10000 * 4),
            "n_units": n_units5, "n_surrogates": 10000, "n_permutations": N_PERM_TOTAL,
            "n_clusters": int(len(out5["cbpa"][("happy", "silence")]["cluster_pv"])),
            "config": "BASELINE config 5 end to end, wall clock: sweep.cmc_surrogate_cbpa_sweep(units, ...) - 80 "
                      "subject-conditions (20 subjects x 4 conditions, 64x64 channels, 30 epochs x 4 s, pinned float32 host "
                      "tensors in, numpy out) dealt round-robin over the ranks, each with a 10,000-surrogate phase null, "
                      "then the 10,000-permutation CBPA of one condition contrast (20 x 100 x 64) sharded over the ranks",
            "api": "multimodal_biosignal_analysis_b200.sweep.cmc_surrogate_cbpa_sweep",
        }
        if world == 1:
            # what a caller holding the reference's arrays pays: pageable numpy float32 / float64 (np.load output) are
            # converted into pinned staging buffers by host threads before the upload
            n_small = 8
            for label, conv in (("numpy_float32_pageable", lambda a: a), ("numpy_float64_pageable", lambda a: a.astype(np.float64))):
                hs = [(conv(a), conv(b)) for a, b in host_sets[:2]]
                us = {(f"S{s_:02d}", c_): hs[(2 * s_ + k_) % 2] for s_ in range(n_small // 2) for k_, c_ in enumerate(conds5[:2])}
                csweep.cmc_surrogate_cbpa_sweep({k: us[k] for k in list(us)[:2]}, FS, nperseg=NPERSEG, freq_band=BAND,
                                                segment_starts=starts_h, n_surrogates=10000, seed=11, contrasts=[])
                t0 = time.perf_counter()
                csweep.cmc_surrogate_cbpa_sweep(us, FS, nperseg=NPERSEG, freq_band=BAND, segment_starts=starts_h,
                                                n_surrogates=10000, seed=11, contrasts=[])
                stages["config5_sweep"][f"ms_per_unit_wall_{label}"] = (time.perf_counter() - t0) * 1e3 / len(us)
                del hs, us

        # ---- stage: BASELINE config 1 (the reference's own CPU-runnable case): one EEG x one bipolar EMG channel ----
        eeg1, emg1 = syn.make_recording(122880, 1, 2, seed=1)
        bip1 = np.ascontiguousarray(emg1[:, :1] - emg1[:, 1:2])
        for _ in range(3):
            c1 = sf.welch_magnitude_squared_coherence(eeg1, bip1, FS, nperseg=1024).coherence
        t0 = time.perf_counter()
        for _ in range(20):
            c1 = sf.welch_magnitude_squared_coherence(eeg1, bip1, FS, nperseg=1024).coherence
        cfg1_ms = (time.perf_counter() - t0) * 1e3 / 20
        stages["config1_welch_pair"] = {
            "metric": "pair_spectra_per_s", "value": 1.0 / (cfg1_ms / 1e3), "unit": "pair-spectra/s", "ms": cfg1_ms,
            "scaling": "replicated",
            "config": "BASELINE config 1: C3 x one bipolar EMG channel, 60 s @ 2048 Hz, nperseg 1024 (L = 239, F = 513): one "
                      "blocking welch_magnitude_squared_coherence call, numpy in, numpy out (launch / PCIe latency bound)"}
        if rank == 0 and world == 1 and not args.no_cpu:
            from scipy import signal as ssig
            x64, y64 = eeg1[:, 0].astype(np.float64), bip1[:, 0].astype(np.float64)
            ssig.coherence(x64, y64, fs=FS, nperseg=1024)
            t0 = time.perf_counter()
            for _ in range(20):
                _, cref1 = ssig.coherence(x64, y64, fs=FS, nperseg=1024)
            sc_ms = (time.perf_counter() - t0) * 1e3 / 20
            stages["config1_welch_pair"]["cpu_baseline"] = {
                "value": 1.0 / (sc_ms / 1e3), "unit": "pair-spectra/s", "cores": 1, "kind": "reference",
                "sample": "scipy.signal.coherence(fs=2048, nperseg=1024) - the reference's own CPU path for this config "
                          "(preprocessing.py:1228-1230) - 20 repetitions, one core",
                "max_abs_diff_vs_gpu": float(np.max(np.abs(c1[:, 0, 0] - cref1)))}

        # ---- stage: the production call of the reference workflow (subject_feature_extraction_workflow.py:58-69) ----
        n_prod = int(FS) * 600
        g = torch.Generator(device=dev).manual_seed(5)
        eeg_p = torch.randn((n_prod, 11), device=dev, generator=g)
        emg_p = torch.randn((n_prod, 64), device=dev, generator=g)

        def prod():
            return sf.multitaper_magnitude_squared_coherence(eeg_p, emg_p, FS, window_length_sec=2.0, use_jackknife=True,
                                                             reduce_emg=True, zero_nonsignificant=True, freq_band=(1, 100))
        for _ in range(2):
            rp = prod()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(5):
            rp = prod()
        torch.cuda.synchronize()
        prod_ms = (time.perf_counter() - t0) * 1e3 / 5
        w_prod = int(rp["coherence_raw"].shape[0])
        stages["production_multitaper"] = {
            "metric": "window_pair_spectra_per_s", "value": w_prod * 11 * 64 * world / (prod_ms / 1e3),
            "unit": "pair-spectra/s (one per window)", "ms": prod_ms, "windows": w_prod, "scaling": "weak",
            "config": "production geometry: 10-minute recording, 11 EEG x 64 EMG channels, N = 4096 / hop 2048, K = 5 DPSS "
                      "tapers, jackknife CI + significance zeroing + EMG-argmax fused (compute_task_wise_aggregated_cmc "
                      "semantics), 1-100 Hz, device-resident input, wall clock of the blocking call"}
        del eeg_p, emg_p, rp

    # ---- CPU baseline (rank 0, N = 1): oracle port on a bounded sample ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        v, ms, cores, sample = run_cpu_reference(steps=2, warmup=1, sample_epochs=10)
        cpu = {"value": v, "unit": "pair-spectra/s", "cores": cores, "kind": "port", "sample": sample,
               "ms_per_step": ms}
        if stages:
            for key, base in run_cpu_stage_baselines().items():   # CPU ports of the surrogate / CBPA stages
                stages[key]["cpu_baseline"] = base

    if rank == 0:
        line = {
            "metric": "pair_spectra_per_s", "value": value, "unit": "pair-spectra/s", "n_gpus": world,
            "steps": steps, "warmup": warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None,
            "dtype": ("bf16x3 -> f32 accumulate (K1t DFT), " if k1_plan is not None else "f32 (FFT), ") +
                     "tf32x3 -> f32 accumulate (CSD)",
            "data": "synthetic", "config": workload_config(),
            "roofline": {"bound": "hbm", "kernel": k1_name,
                         "achieved": k1_gbs, "peak": hbm, "unit": "GB/s", "frac": k1_gbs / hbm,
                         # dram__bytes_read + dram__bytes_write of one K1 launch, ncu --set full (profiles/r01b_k1_tma.md,
                         # addendum 8: 63.0 MB read + 4.0 MB written inside the window, the rest of the output still in L2)
                         "traffic": K1_TRAFFIC_PROFILED["bytes"], "traffic_source": K1_TRAFFIC_PROFILED["source"],
                         "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": int(k1_bytes), "launch_ms": k1_ms,
                         "k2_ms_per_step": k2_ms, "k1_share_of_step": k1_ms / (k1_ms + k2_ms),
                         "timing_note": "launch_ms / k2_ms_per_step: CUDA events around graphs of 8 back-to-back launches "
                                        "of each kernel on one stream right after the timed region (serial step "
                                        f"{serial_ms_per_step:.4f} ms); the timed region replays {n_big} graphs of {G} "
                                        f"pipelined steps + {rem} single steps, in which kernels of neighbouring steps "
                                        f"overlap"},
            "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks, "stages": stages,
            "host_enqueue_ms_per_step": host_enqueue_ms,
            # the other two headline metrics of BASELINE.json, copied up from `stages` for convenience
            "surrogates_per_s": stages.get("surrogate_null_phase", {}).get("value"),
            "cbpa_permutations_per_s": stages.get("cbpa", {}).get("value"),
            "config5_sweep_seconds": stages.get("config5_sweep", {}).get("seconds_total"),
        }
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=500)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline leg")
    ap.add_argument("--graph-steps", type=int, default=20,
                    help="steps per CUDA graph in the timed region (0: one graph pair per step)")
    ap.add_argument("--k1", default="tc", choices=["tc", "fft"],
                    help="K1 of the headline step: tensor-core half-block DFT (default) or the FFT kernel")
    ap.add_argument("--skip-stages", action="store_true",
                    help="headline metric only (profiling runs): no surrogate / CBPA / sweep stages")
    args = ap.parse_args()
    if args.impl == "reference":
        main_reference(args)
    else:
        main_gpu(args)


if __name__ == "__main__":
    main()
