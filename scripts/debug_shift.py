import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from scipy import signal
from multimodal_biosignal_analysis_b200 import kernels as K, synthetic as syn

def run(shift_vals, label):
    N, hop, ep, n_epochs = 512, 256, 2048, 6
    eeg, emg = syn.make_epochs(n_epochs, ep, 20, 70, seed=11)
    starts = syn.epoch_segment_starts(n_epochs, ep, N, hop)
    win = torch.as_tensor(signal.get_window("hann", N).astype(np.float32)[None]).cuda()
    X = K.fft_segments(torch.as_tensor(eeg).cuda(), torch.as_tensor(starts).cuda(), win, 1, 1, 40)[:, 0]
    Y = K.fft_segments(torch.as_tensor(emg).cuda(), torch.as_tensor(starts).cuda(), win, 1, 1, 40)[:, 0]
    res = K.csd_msc(X, Y)
    torch.cuda.synchronize()
    sh = torch.as_tensor(np.asarray(shift_vals, dtype=np.int32)).cuda()
    ex, ms = K.surrogate_null(res, K.SURR_SHIFT, 0, len(shift_vals), shifts=sh)
    torch.cuda.synchronize()
    print(label, "ok", ms.cpu().numpy()[:4], int(ex.sum()), flush=True)

which = sys.argv[1]
if which == "even": run([2, 4, 8, 2], "even")
elif which == "odd": run([1, 3, 5, 1], "odd")
elif which == "one": run([4], "one")
