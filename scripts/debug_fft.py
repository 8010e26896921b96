import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from scipy import signal
from multimodal_biosignal_analysis_b200 import kernels as K
case = sys.argv[1]
rng = np.random.default_rng(3)
N = 2048
nch = {"a": 64, "b": 11, "c": 64, "d": 64, "e": 64}[case]
eeg = torch.as_tensor(rng.standard_normal((3 * N, nch)).astype(np.float32)).cuda()
nseg = 5 if case != "e" else 4
starts = torch.as_tensor((np.arange(nseg) * (N // 2)).astype(np.int64)).cuda()
win = torch.as_tensor(signal.get_window("hann", N).astype(np.float32)[None]).cuda()
if case in ("a", "b", "e"):
    out = torch.zeros((nseg, 1, 100, 128), dtype=torch.complex64, device="cuda")
    K.fft_segments(eeg, starts, win, 1, 1, 100, out=out, ch_offset=0)
elif case == "c":
    out = K.fft_segments(eeg, starts, win, 1, 1, 100)
elif case == "d":
    out = K.fft_segments(eeg, starts, win, 1, 0, 1024)
torch.cuda.synchronize()
print(case, "ok", float(out.abs().sum()))
