"""Production geometry of the reference workflow (subject_feature_extraction_workflow.py:58-69): 11 EEG channels x
64 HD-EMG channels, 2 s windows (N = 4096) with 50 % overlap, K = 5 tapers, jackknife CI, EMG-argmax reduction,
whole recording of 10 minutes at 2048 Hz."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from multimodal_biosignal_analysis_b200 import signal_features as sf
rng = np.random.default_rng(0)
n = 2048 * 600
eeg = torch.from_numpy(rng.standard_normal((n, 11)).astype(np.float32)).cuda()
emg = torch.from_numpy(rng.standard_normal((n, 64)).astype(np.float32)).cuda()
for full in (False, True):
    for _ in range(2):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        r = sf.multitaper_magnitude_squared_coherence(eeg, emg, 2048.0, window_length_sec=2.0, use_jackknife=True,
                                                      reduce_emg=True, zero_nonsignificant=True,
                                                      freq_band=None if full else (1, 100))
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
    c = r["coherence_raw"]
    print("all bins" if full else "1-100 Hz", tuple(c.shape), f"{dt * 1e3:.1f} ms for {c.shape[0]} windows")

# where does the in-band call spend its time?  kernels (CUDA events) vs wall clock, and a cProfile of the host side
import cProfile, pstats, io
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); t0 = time.perf_counter(); e0.record()
r = sf.multitaper_magnitude_squared_coherence(eeg, emg, 2048.0, window_length_sec=2.0, use_jackknife=True,
                                              reduce_emg=True, zero_nonsignificant=True, freq_band=(1, 100))
e1.record(); t_host = time.perf_counter() - t0; torch.cuda.synchronize(); t_all = time.perf_counter() - t0
print(f"host returns after {t_host * 1e3:.1f} ms, device done after {t_all * 1e3:.1f} ms, GPU span {e0.elapsed_time(e1):.1f} ms")
pr = cProfile.Profile(); pr.enable()
r = sf.multitaper_magnitude_squared_coherence(eeg, emg, 2048.0, window_length_sec=2.0, use_jackknife=True,
                                              reduce_emg=True, zero_nonsignificant=True, freq_band=(1, 100))
torch.cuda.synchronize(); pr.disable()
st = io.StringIO(); pstats.Stats(pr, stream=st).sort_stats("cumulative").print_stats(14); print(st.getvalue()[-2600:])
