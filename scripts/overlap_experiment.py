"""Experiment: K2 of recording i on a second stream while K1 of recording i + 1 runs (per-slot spectra buffers)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from scipy import signal
from multimodal_biosignal_analysis_b200 import kernels as K, synthetic as syn
dev = torch.device("cuda")
NR, NE, NM = 4, 64, 64
starts = torch.from_numpy(syn.epoch_segment_starts(30, 8192, 2048, 1024)).to(dev)
win = torch.from_numpy(signal.get_window("hann", 2048).astype(np.float32)[None]).to(dev)
L = len(starts)
sets = []
for r in range(NR):
    eeg, emg = syn.make_epochs(30, 8192, NE, NM, seed=r)
    sets.append((torch.from_numpy(eeg).to(dev), torch.from_numpy(emg).to(dev),
                 torch.empty((L, 1, 100, NE + NM), dtype=torch.complex64, device=dev)))
for e, m, sp in sets:
    K.fft_segments(e, starts, win, 1, 1, 100, out=sp, ch_offset=0)
    K.fft_segments(m, starts, win, 1, 1, 100, out=sp, ch_offset=NE)
    K.csd_msc(sp[:, 0, :, :NE], sp[:, 0, :, NE:])
torch.cuda.synchronize()
FORK = os.environ.get("FORK_K1") is not None      # the two K1 launches as parallel branches of the graph
side = torch.cuda.Stream()
graphs = []
for e, m, sp in sets:
    gA, gB = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
    with torch.cuda.graph(gA):
        if FORK:
            side.wait_stream(torch.cuda.current_stream())
            K.fft_segments(e, starts, win, 1, 1, 100, out=sp, ch_offset=0)
            with torch.cuda.stream(side):
                K.fft_segments(m, starts, win, 1, 1, 100, out=sp, ch_offset=NE)
            torch.cuda.current_stream().wait_stream(side)
        else:
            K.fft_segments(e, starts, win, 1, 1, 100, out=sp, ch_offset=0)
            K.fft_segments(m, starts, win, 1, 1, 100, out=sp, ch_offset=NE)
    with torch.cuda.graph(gB):
        res = K.csd_msc(sp[:, 0, :, :NE], sp[:, 0, :, NE:])
    graphs.append((gA, gB, res))
sA, sB = torch.cuda.Stream(), torch.cuda.Stream()
sA2 = torch.cuda.Stream()
ALT = os.environ.get("ALT_K1") is not None        # K1 graphs of even / odd steps on two streams
steps = 400

def run(overlap):
    evA = [torch.cuda.Event() for _ in range(steps)]
    evB = [torch.cuda.Event() for _ in range(steps)]
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    with torch.cuda.stream(sA):
        t0.record()
    for i in range(steps):
        gA, gB, _ = graphs[i % NR]
        sK1 = sA2 if (ALT and overlap and (i & 1)) else sA
        with torch.cuda.stream(sK1):
            if overlap and i >= NR:
                sK1.wait_event(evB[i - NR])         # the slot's spectra are free again
            gA.replay()
            evA[i].record()
        with torch.cuda.stream(sB if overlap else sA):
            if overlap:
                sB.wait_event(evA[i])
            gB.replay()
            evB[i].record()
    with torch.cuda.stream(sB if overlap else sA):
        t1.record()
    torch.cuda.synchronize()
    return t0.elapsed_time(t1) / steps

for _ in range(2):
    print("serial  ms/step", run(False), " overlapped ms/step", run(True))
