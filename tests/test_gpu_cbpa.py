"""GPU parity: K4 (CBPA) vs the MNE-algorithm oracle - t-map, label map, fixed-point masses, H0
and p-values bit-exact for a host-supplied sign table."""
import numpy as np
import pytest
import torch
from scipy.stats import t as t_dist

from oracle import cbpa as ocb
from multimodal_biosignal_analysis_b200 import synthetic as syn

pytestmark = pytest.mark.gpu


def _csr(adj):
    a = adj.tocsr()
    a.sort_indices()
    return (torch.as_tensor(a.indptr.astype(np.int32)).cuda(), torch.as_tensor(a.indices.astype(np.int32)).cuda())


def _run(X, signs, thr, tail, adj):
    from multimodal_biosignal_analysis_b200 import kernels as K
    n_subj = X.shape[0]
    Xf = np.ascontiguousarray(X.reshape(n_subj, -1))
    indptr, indices = _csr(adj)
    Xd = torch.as_tensor(Xf).cuda()
    t_obs, labels, mass, n = K.cbpa_observed(Xd, thr, tail, indptr, indices)
    h0 = K.cbpa_permute(Xd, torch.as_tensor(signs).cuda(), 0, len(signs), thr, tail, indptr, indices)
    return t_obs.cpu().numpy(), labels.cpu().numpy(), mass.cpu().numpy(), n, h0.cpu().numpy()


@pytest.mark.parametrize("tail", [0, 1, -1])
def test_cbpa_cfg4_small_permutation_set_bit_exact(cuda_device, tail):
    pos = syn.sensor_positions(64)
    adj = ocb.combine_adjacency(100, ocb.delaunay_adjacency(pos))
    X = syn.make_cbpa_contrast(20, 100, 64)
    if tail == -1:
        X = -X
    signs = syn.make_sign_table(48, 20, seed=42, tail=tail)
    thr = t_dist.ppf(0.975 if tail == 0 else 0.95, 19)
    if tail == -1:
        thr = -thr
    ref = ocb.permutation_cluster_1samp_test(X, signs, thr, tail, adj)
    t_obs, labels, mass, n, h0 = _run(X, signs, thr, tail, adj)
    np.testing.assert_array_equal(t_obs, ref["t_obs"].reshape(-1))       # fp64 bit-exact
    assert n == len(ref["clusters"])
    np.testing.assert_array_equal(labels, ref["labels"])
    np.testing.assert_array_equal(mass, ref["mass_fixed"])
    np.testing.assert_array_equal(h0, ref["H0_fixed"][1:])
    h0_full = np.concatenate([[ref["H0_fixed"][0]], h0])
    np.testing.assert_array_equal(ocb.pvalues_from_h0(mass, h0_full, tail), ref["cluster_pv"])


def test_cbpa_production_shape_with_phase_wrap_and_nan(cuda_device):
    rng = np.random.default_rng(9)
    n_subj, n_times, n_ch = 13, 36, 11
    pos = syn.sensor_positions(64)[:n_ch]
    adj = ocb.add_phase_wraparound(ocb.combine_adjacency(n_times, ocb.delaunay_adjacency(pos)), n_times, n_ch)
    X = rng.standard_normal((n_subj, n_times, n_ch))
    X[:, 34:, :4] += 4.0
    X[:, :2, :4] += 4.0          # cluster that only connects through the wrap-around edges
    X[:, 10, 3] = np.nan          # all-NaN bin: t is NaN and must never enter a cluster
    signs = syn.make_sign_table(64, n_subj, seed=1)
    thr = t_dist.ppf(0.975, n_subj - 1)
    ref = ocb.permutation_cluster_1samp_test(X, signs, thr, 0, adj)
    t_obs, labels, mass, n, h0 = _run(X, signs, thr, 0, adj)
    np.testing.assert_array_equal(np.isnan(t_obs), np.isnan(ref["t_obs"].reshape(-1)))
    ok = ~np.isnan(t_obs)
    np.testing.assert_array_equal(t_obs[ok], ref["t_obs"].reshape(-1)[ok])
    np.testing.assert_array_equal(labels, ref["labels"])
    np.testing.assert_array_equal(mass, ref["mass_fixed"])
    np.testing.assert_array_equal(h0, ref["H0_fixed"][1:])
    wrap_lab = labels.reshape(n_times, n_ch)
    assert wrap_lab[35, 0] != 0 and wrap_lab[35, 0] == wrap_lab[0, 0]


def test_cbpa_sign_symmetry_and_no_cluster(cuda_device):
    """flipping every subject maps t -> -t and H0 -> -H0; a huge threshold gives H0 = 0."""
    pos = syn.sensor_positions(64)[:16]
    adj = ocb.combine_adjacency(12, ocb.delaunay_adjacency(pos))
    X = syn.make_cbpa_contrast(10, 12, 16, seed=3)
    signs = syn.make_sign_table(32, 10, seed=5)
    thr = t_dist.ppf(0.975, 9)
    _, _, _, _, h0 = _run(X, signs, thr, 0, adj)
    _, _, _, _, h0_neg = _run(X, (-signs).astype(np.int8), thr, 0, adj)
    np.testing.assert_array_equal(h0, -h0_neg)
    t_obs, labels, mass, n, h0_big = _run(X, signs, 1e6, 0, adj)
    assert n == 0 and np.all(labels == 0) and np.all(h0_big == 0)


def test_cbpa_full_cfg4_all_permutations_vs_oracle(cuda_device):
    """BASELINE config 4 at full size: ALL 1,024 sign-flip permutations against the oracle (bit-exact fixed-point
    H0, ~3 s of numpy / scipy), plus sharding invariance of the permutation range."""
    from multimodal_biosignal_analysis_b200 import kernels as K
    pos = syn.sensor_positions(64)
    adj = ocb.combine_adjacency(100, ocb.delaunay_adjacency(pos))
    X = syn.make_cbpa_contrast(20, 100, 64)
    signs = syn.make_sign_table(1024, 20, seed=42)
    thr = t_dist.ppf(0.975, 19)
    indptr, indices = _csr(adj)
    Xd = torch.as_tensor(np.ascontiguousarray(X.reshape(20, -1))).cuda()
    sd = torch.as_tensor(signs).cuda()
    whole = K.cbpa_permute(Xd, sd, 0, 1024, thr, 0, indptr, indices).cpu().numpy()
    parts = np.concatenate([K.cbpa_permute(Xd, sd, a, a + 128, thr, 0, indptr, indices).cpu().numpy()
                            for a in range(0, 1024, 128)])
    np.testing.assert_array_equal(whole, parts)
    with np.errstate(all="ignore"):
        ref = ocb.permutation_cluster_1samp_test(X, signs, thr, 0, adj)
    np.testing.assert_array_equal(whole, ref["H0_fixed"][1:])
    t_obs, labels, mass, n = K.cbpa_observed(Xd, thr, 0, indptr, indices)
    np.testing.assert_array_equal(labels.cpu().numpy(), ref["labels"])
    np.testing.assert_array_equal(mass.cpu().numpy(), ref["mass_fixed"])


def test_cbpa_max_map_size_supra_list_overflow_and_compact_paths(cuda_device):
    """n_tests = 16384 (the ABI maximum): the supra-threshold list holds only part of the map, so a strong
    effect (most nodes supra-threshold) takes the whole-map fallback and a null map takes the compact path;
    both must match the oracle bit for bit."""
    rng = np.random.default_rng(31)
    n_subj, n_times, n_ch = 9, 256, 64
    adj = ocb.combine_adjacency(n_times, ocb.delaunay_adjacency(syn.sensor_positions(n_ch)))
    signs = syn.make_sign_table(6, n_subj, seed=3)
    all_minus = -np.ones((1, n_subj), dtype=signs.dtype)        # as many supra-threshold nodes as the observed map
    one_flip = np.ones((1, n_subj), dtype=signs.dtype)
    one_flip[0, 4] = -1
    signs = np.concatenate([signs[:3], all_minus, signs[3:], one_flip])
    thr = t_dist.ppf(0.975, n_subj - 1)
    for shift in (0.0, 1.5):
        X = rng.standard_normal((n_subj, n_times, n_ch)) + shift
        ref = ocb.permutation_cluster_1samp_test(X, signs, thr, 0, adj)
        t_obs, labels, mass, n, h0 = _run(X, signs, thr, 0, adj)
        np.testing.assert_array_equal(t_obs, ref["t_obs"].reshape(-1))
        np.testing.assert_array_equal(labels, ref["labels"])
        np.testing.assert_array_equal(mass, ref["mass_fixed"])
        np.testing.assert_array_equal(h0, ref["H0_fixed"][1:])


@pytest.mark.parametrize("n_subj,n_times,n_ch", [(2, 10, 7), (3, 9, 5), (32, 12, 6), (35, 11, 9), (21, 33, 1)])
def test_cbpa_subject_counts_and_ragged_maps(cuda_device, n_subj, n_times, n_ch):
    """Smallest, template-boundary (32) and generic (> 32) subject counts, maps whose size is not a multiple of
    the 32-test tile of the re-laid data, exact zeros / constant columns (t = NaN) and huge / tiny magnitudes that
    leave the fast constant-division band: everything bit-identical to the oracle."""
    rng = np.random.default_rng(100 + n_subj)
    pos = syn.sensor_positions(64)[:max(n_ch, 3)]
    sp = ocb.delaunay_adjacency(pos)[:n_ch][:, :n_ch] if n_ch >= 3 else None
    if sp is None:
        from scipy import sparse
        sp = sparse.csr_matrix((n_ch, n_ch))
    adj = ocb.combine_adjacency(n_times, sp)
    X = rng.standard_normal((n_subj, n_times, n_ch)) + 0.6
    X[:, 0, 0] = 0.0                    # all-zero column: t = 0/0
    X[:, 1, 0] = 3.25                   # constant column: zero variance
    X[:, 2, 0] *= 1e200                 # outside the fast-division exponent band
    X[:, 3, 0] *= 1e-200
    signs = syn.make_sign_table(24, n_subj, seed=n_subj)
    thr = t_dist.ppf(0.975, n_subj - 1)
    with np.errstate(all="ignore"):
        ref = ocb.permutation_cluster_1samp_test(X, signs, thr, 0, adj)
    t_obs, labels, mass, n, h0 = _run(X, signs, thr, 0, adj)
    np.testing.assert_array_equal(t_obs, ref["t_obs"].reshape(-1))       # NaNs compare equal position-wise
    assert n == len(ref["clusters"])
    np.testing.assert_array_equal(labels, ref["labels"])
    np.testing.assert_array_equal(mass, ref["mass_fixed"])
    np.testing.assert_array_equal(h0, ref["H0_fixed"][1:])


def test_cbpa_lockfree_labelling_is_deterministic_under_repetition(cuda_device):
    """compute-sanitizer (racecheck) is closed on the GPU pool, so the lock-free shared-memory union-find
    (csrc/cbpa.cu: atomicCAS hooks, atomic fixed-point masses, permutations claimed from a device counter) is
    checked the other way round: a race would show as run-to-run variation, so 40 repetitions of the same 1,500
    permutations - whole range, ragged sub-ranges, shared and private workspaces - must agree bit for bit with each
    other and with the oracle on a sample."""
    from multimodal_biosignal_analysis_b200 import kernels as K
    rng = np.random.default_rng(99)
    n_subj, n_times, n_ch = 14, 60, 64
    adj = ocb.combine_adjacency(n_times, ocb.delaunay_adjacency(syn.sensor_positions(n_ch)))
    X = rng.standard_normal((n_subj, n_times, n_ch)) + 0.35          # many supra-threshold nodes, large clusters
    signs = syn.make_sign_table(1500, n_subj, seed=5)
    thr = t_dist.ppf(0.975, n_subj - 1)
    indptr, indices = _csr(adj)
    Xd = torch.as_tensor(np.ascontiguousarray(X.reshape(n_subj, -1))).cuda()
    sd = torch.as_tensor(signs).cuda()
    ws = K.cbpa_workspace(Xd)
    K.cbpa_observed(Xd, thr, 0, indptr, indices, ws=ws)
    first = K.cbpa_permute(Xd, sd, 0, 1500, thr, 0, indptr, indices, ws=ws, tiled=True).cpu().numpy()
    for rep in range(40):
        if rep % 3 == 0:
            got = K.cbpa_permute(Xd, sd, 0, 1500, thr, 0, indptr, indices, ws=ws, tiled=True).cpu().numpy()
        elif rep % 3 == 1:
            got = K.cbpa_permute(Xd, sd, 0, 1500, thr, 0, indptr, indices).cpu().numpy()
        else:
            cut = int(rng.integers(1, 1499))
            got = np.concatenate([K.cbpa_permute(Xd, sd, 0, cut, thr, 0, indptr, indices).cpu().numpy(),
                                  K.cbpa_permute(Xd, sd, cut, 1500, thr, 0, indptr, indices).cpu().numpy()])
        np.testing.assert_array_equal(got, first)
    pick = np.arange(0, 1500, 100)
    with np.errstate(all="ignore"):
        ref = ocb.permutation_cluster_1samp_test(X, signs[pick], thr, 0, adj)
    np.testing.assert_array_equal(first[pick], ref["H0_fixed"][1:])


def test_empty_permutation_shard(cuda_device):
    """A rank whose slice of a short (exact-enumeration) sign table is empty - 2 subjects have ONE sign pattern, two
    ranks share it - must get an empty H0 slice back, not an error (found by the 2-GPU bench of config 5)."""
    from multimodal_biosignal_analysis_b200 import cbpa as cb, kernels as K, synthetic as syn
    X = syn.make_cbpa_contrast(2, 6, 16, seed=4)
    signs = cb.make_sign_table(1000, 2, seed=1, tail=0)
    assert signs.shape == (1, 2)
    adj = cb.combine_adjacency(6, cb.find_ch_adjacency_from_positions(syn.sensor_positions(64)[:16])).tocsr()
    adj.sort_indices()
    Xd = torch.from_numpy(np.ascontiguousarray(X.reshape(2, -1))).cuda()
    ip = torch.from_numpy(adj.indptr.astype(np.int32)).cuda()
    ix = torch.from_numpy(adj.indices.astype(np.int32)).cuda()
    sd = torch.from_numpy(signs).cuda()
    h_empty = K.cbpa_permute(Xd, sd, 1, 1, 12.7, 0, ip, ix)
    h_one = K.cbpa_permute(Xd, sd, 0, 1, 12.7, 0, ip, ix)
    assert h_empty.numel() == 0 and h_one.numel() == 1
