"""torchrun check: a surrogate null / CBPA shared by several GPUs returns exactly what one GPU returns.
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 scripts/check_multi_gpu.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from scipy.stats import t as t_dist
from multimodal_biosignal_analysis_b200 import signal_features as sf, data_surrogation as ds, cbpa as cb, synthetic as syn
from multimodal_biosignal_analysis_b200 import kernels as K, dist as cd

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl")
eeg, emg = syn.make_epochs(6, 4096, 12, 20, seed=77)
starts = syn.epoch_segment_starts(6, 4096, 1024, 512)
pc = sf.welch_magnitude_squared_coherence(eeg, emg, 2048.0, nperseg=1024, freq_band=(2, 60), segment_starts=starts)
csd = pc.device_result
F = csd.dims[1]
n = 300
# single-GPU references (every rank computes them locally, no collectives)
ex1, ms1 = K.surrogate_null(csd, K.SURR_PHASE, 0, n, seed=5)
shifts = np.random.default_rng(2).integers(1, len(starts), n).astype(np.int32)
ex2, ms2 = K.surrogate_null(csd, K.SURR_SHIFT, 0, n, shifts=torch.from_numpy(shifts).cuda())
ok = True
for shard in ("frequency", "surrogate"):
    r = ds.phase_randomised_surrogate_null(pc, n, seed=5, shard=shard)
    ok &= np.array_equal(r["exceed"], ex1.cpu().numpy()) and np.array_equal(r["max_stat"], ms1.cpu().numpy())
    r = ds.circular_shift_surrogate_null(pc, n, shifts=shifts, shard=shard)
    ok &= np.array_equal(r["exceed"], ex2.cpu().numpy()) and np.array_equal(r["max_stat"], ms2.cpu().numpy())
# CBPA: permutations sharded by index, H0 all-gathered
X = syn.make_cbpa_contrast(12, 20, 16, seed=2)
adj = cb.combine_adjacency(20, cb.find_ch_adjacency_from_positions(syn.sensor_positions(64)[:16]))
signs = syn.make_sign_table(257, 12, seed=3)
thr = float(t_dist.ppf(0.975, 11))
Xd = torch.from_numpy(np.ascontiguousarray(X.reshape(12, -1))).cuda()
a = adj.tocsr(); a.sort_indices()
ip, ix = torch.from_numpy(a.indptr.astype(np.int32)).cuda(), torch.from_numpy(a.indices.astype(np.int32)).cuda()
sd = torch.from_numpy(signs).cuda()
h_all = K.cbpa_permute(Xd, sd, 0, len(signs), thr, 0, ip, ix)
b, e = cd.shard_range(len(signs))
h_sh = cd.all_gather_ranges(K.cbpa_permute(Xd, sd, b, e, thr, 0, ip, ix), len(signs))
ok &= bool(torch.equal(h_all, h_sh))
# per-pair thresholds: ranks split the frequency axis of the histogram passes
thr1, _ = ds.null_quantile_thresholds(csd, n, 5, 0.95, passes=2, shard="surrogate" if world == 1 else "auto")
ref_thr = None
if True:
    # single-GPU reference without collectives: run the passes on the whole axis by hand
    import multimodal_biosignal_analysis_b200.dist as _cd
    saved = _cd.world
    _cd.world = lambda: (0, 1)
    try:
        ref_thr, _ = ds.null_quantile_thresholds(csd, n, 5, 0.95, passes=2)
    finally:
        _cd.world = saved
ok &= bool(torch.equal(thr1, ref_thr))
# subject-condition sweep: units dealt round-robin, maps all-reduced, CBPA sharded == the same call on one rank
from multimodal_biosignal_analysis_b200 import sweep
units = {}
for s_ in range(5):
    for c_ in ("A", "B"):
        units[(s_, c_)] = syn.make_epochs(3, 2048, 8, 6, seed=500 + 10 * s_ + (c_ == "B"))
adj_sp = cb.find_ch_adjacency_from_positions(syn.sensor_positions(64)[:8])
kw = dict(nperseg=512, freq_band=(4, 60), n_surrogates=64, seed=9, spatial_adjacency=adj_sp, n_permutations=120)
multi = sweep.cmc_surrogate_cbpa_sweep(units, 1024.0, **kw)
_cd.world = lambda: (0, 1)
try:
    single = sweep.cmc_surrogate_cbpa_sweep(units, 1024.0, **kw)
finally:
    _cd.world = saved
ok &= np.array_equal(multi["cmc"], single["cmc"]) and np.array_equal(multi["n_significant_pairs"], single["n_significant_pairs"])
ok &= np.allclose(multi["threshold_fwe"], single["threshold_fwe"], rtol=1e-6)
ra, rb = multi["cbpa"][("A", "B")], single["cbpa"][("A", "B")]
ok &= np.array_equal(ra["H0"], rb["H0"]) and np.array_equal(ra["cluster_pv"], rb["cluster_pv"]) and np.array_equal(ra["t_obs"], rb["t_obs"])
flag = torch.tensor([1 if ok else 0], device="cuda")
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print(f"world={world}: sharded nulls and CBPA identical to one GPU: {bool(flag.item())}")
dist.destroy_process_group()
sys.exit(0 if flag.item() else 1)
