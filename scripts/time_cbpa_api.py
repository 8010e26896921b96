"""End-to-end timing of the MNE-signature CBPA front-end at the config 4 geometry (host in, host out)."""
import sys, os, time, cProfile, pstats, io
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from scipy.stats import t as t_dist
from multimodal_biosignal_analysis_b200 import cbpa as cb, synthetic as syn
X = syn.make_cbpa_contrast(20, 100, 64)
sp = cb.find_ch_adjacency_from_positions(syn.sensor_positions(64))
adj = cb.combine_adjacency(100, sp)
thr = float(t_dist.ppf(0.975, 19))
for n_perm in (1024, 10000):
    for rep in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        t_obs, clusters, pv, H0 = cb.permutation_cluster_1samp_test(X, threshold=thr, n_permutations=n_perm, tail=0,
                                                                   adjacency=adj, out_type="mask", seed=42)
        dt = time.perf_counter() - t0
    print(f"n_perm={n_perm}: {dt * 1e3:.1f} ms end to end ({n_perm / dt / 1e6:.2f} M permutations/s), {len(clusters)} clusters")
pr = cProfile.Profile(); pr.enable()
cb.permutation_cluster_1samp_test(X, threshold=thr, n_permutations=10000, tail=0, adjacency=adj, out_type="mask", seed=42)
pr.disable()
st = io.StringIO(); pstats.Stats(pr, stream=st).sort_stats("cumulative").print_stats(16); print(st.getvalue()[-3000:])
