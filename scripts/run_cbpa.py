import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from scipy.stats import t as t_dist
from multimodal_biosignal_analysis_b200 import kernels as K, synthetic as syn
from multimodal_biosignal_analysis_b200.cbpa import combine_adjacency, find_ch_adjacency_from_positions
adj = combine_adjacency(100, find_ch_adjacency_from_positions(syn.sensor_positions(64))); adj.sort_indices()
X = torch.from_numpy(np.ascontiguousarray(syn.make_cbpa_contrast(20, 100, 64).reshape(20, -1))).cuda()
signs = torch.from_numpy(syn.make_sign_table(10000, 20, seed=42)).cuda()
ip = torch.from_numpy(adj.indptr.astype(np.int32)).cuda(); ix = torch.from_numpy(adj.indices.astype(np.int32)).cuda()
thr = float(t_dist.ppf(0.975, 19))
for _ in range(3):
    h0 = K.cbpa_permute(X, signs, 0, 10000, thr, 0, ip, ix)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    h0 = K.cbpa_permute(X, signs, 0, 10000, thr, 0, ip, ix)
e1.record(); torch.cuda.synchronize()
print("ok", int(h0.abs().max()), "ms/10000 perms", e0.elapsed_time(e1) / 10, "checksum", int(h0.sum()))
import ctypes
from multimodal_biosignal_analysis_b200 import _lib
lib = _lib.load()
if hasattr(lib, "cmc_dbg_cbpa_cycles"):
    buf = (ctypes.c_ulonglong * 8)()
    lib.cmc_dbg_cbpa_cycles(buf, 1)
    h0 = K.cbpa_permute(X, signs, 0, 10000, thr, 0, ip, ix); torch.cuda.synchronize()
    lib.cmc_dbg_cbpa_cycles(buf, 0)
    tot = sum(buf)
    print("phase cycles per perm [signs, tmap, hook, mass, max]:", [round(b / 10000) for b in buf[:5]], "share", [round(b / tot, 3) for b in buf[:5]])
