"""K4 against MNE-Python itself.  Skipped until ``scripts/make_mne_golden.py`` has been run in an environment with
mne (none is available offline: oracle/cbpa.py is "parity unpinned" until then)."""
import os

import numpy as np
import pytest
from scipy import sparse

from conftest import GOLDEN

PATH = os.path.join(GOLDEN, "mne_cbpa.npz")
pytestmark = pytest.mark.skipif(not os.path.exists(PATH), reason="tests/golden/mne_cbpa.npz not generated (needs mne)")
CASES = ("cfg4_small", "production", "one_tailed")


def _case(g, name):
    X = g[f"{name}_X"]
    n = X.shape[1] * X.shape[2]
    adj = sparse.coo_matrix((np.ones(len(g[f"{name}_adj_row"])), (g[f"{name}_adj_row"], g[f"{name}_adj_col"])), shape=(n, n))
    return X, adj.tocsr(), float(g[f"{name}_thr"]), int(g[f"{name}_tail"])


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_mne(name):
    """oracle/cbpa.py (the restatement every GPU test is checked against) == MNE on MNE's own outputs."""
    from oracle import cbpa as ocb
    from multimodal_biosignal_analysis_b200 import cbpa as cb
    g = np.load(PATH, allow_pickle=False)
    X, adj, thr, tail = _case(g, name)
    ours = cb.combine_adjacency(X.shape[1], sparse.csr_matrix(g[f"{name}_spatial"]))
    if bool(g[f"{name}_wrap"]):
        ours = cb._add_phase_wraparound(ours, X.shape[1], X.shape[2], np.arange(X.shape[1]))
    assert (ours.astype(bool) != adj.astype(bool)).nnz == 0                  # combine_adjacency (+ wrap) identical
    signs = g[f"{name}_signs"] if f"{name}_signs" in g.files else np.ones((0, X.shape[0]), np.int8)
    with np.errstate(all="ignore"):
        ref = ocb.permutation_cluster_1samp_test(X, signs, thr, tail, adj)
    np.testing.assert_allclose(ref["t_obs"], g[f"{name}_t_obs"], rtol=1e-12, atol=1e-12)
    assert len(ref["clusters"]) == len(g[f"{name}_clusters"])
    for a, b in zip(ref["clusters"], g[f"{name}_clusters"]):
        np.testing.assert_array_equal(a, b)                                  # same clusters in the same order
    if len(signs):
        np.testing.assert_allclose(ref["H0"], g[f"{name}_H0"], rtol=1e-9, atol=1e-9)
        np.testing.assert_array_equal(ref["cluster_pv"], g[f"{name}_pv"])


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_gpu_matches_mne(cuda_device, name):
    from multimodal_biosignal_analysis_b200 import cbpa as cb
    g = np.load(PATH, allow_pickle=False)
    if f"{name}_signs" not in g.files:
        pytest.skip("MNE's sign table was not captured")
    X, adj, thr, tail = _case(g, name)
    t_obs, clusters, pv, H0 = cb.spatio_temporal_cluster_1samp_test(X, threshold=thr, tail=tail, adjacency=adj,
                                                                    out_type="mask", signs=g[f"{name}_signs"])
    np.testing.assert_allclose(t_obs, g[f"{name}_t_obs"], rtol=1e-12, atol=1e-12)
    for a, b in zip(clusters, g[f"{name}_clusters"]):
        np.testing.assert_array_equal(a, b)
    np.testing.assert_allclose(H0, g[f"{name}_H0"], rtol=1e-8, atol=1e-8)    # fixed point 2^-30 vs float sums
    np.testing.assert_array_equal(pv, g[f"{name}_pv"])
