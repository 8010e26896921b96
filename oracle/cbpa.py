"""numpy/scipy restatement of the cluster-based permutation test the reference
delegates to MNE-Python (``src/pipeline/cbpa.py:1027-1042``).

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  PARITY UNPINNED at the MNE
boundary: mne is an unpinned dependency (``environment.yml:11``) that is neither
vendored nor installed, and no reference test touches ``cbpa.py``.  This file
restates the published algorithm of ``mne.stats.permutation_cluster_1samp_test``
(``mne/stats/cluster_level.py``: ``_find_clusters``, ``_get_components``,
``_do_1samp_permutations``, ``_pval_from_histogram``; ``mne/stats/parametric.py``:
``ttest_1samp_no_p``) for the options the reference uses: sparse adjacency,
t_power=1, no TFCE, no step-down, out_type='mask'.  Cross-checked against an
independent scipy implementation (``scipy.stats.ttest_1samp`` +
``scipy.ndimage.label``) in ``tests/test_oracle_golden.py``.

Determinism contract shared with the CUDA path:
  * the sign-flip table is HOST-SUPPLIED (int8 in {-1,+1}, one row per
    permutation) - never MNE's / numpy's RNG stream;
  * t-values are fp64 with numpy's operation order: sequential sums over the
    subject axis, mean = sum / n, var = sum((x - mean)^2) / (n - 1),
    t = mean / sqrt(var / n);
  * cluster masses are order-independent int64 fixed-point sums
    ``sum(rint(clip(t, +-T_CLAMP) * 2**30))``; the float64 ``np.sum(t[c])`` MNE
    would compute is reported alongside.
"""
from __future__ import annotations

import numpy as np
from scipy import sparse
from scipy.sparse.csgraph import connected_components

FIX_SHIFT = 30
FIX_SCALE = float(1 << FIX_SHIFT)
T_CLAMP = 65536.0


# --------------------------------------------------------------------------
# adjacency  (cbpa.py:224-243, :949-982; mne.stats.combine_adjacency)
# --------------------------------------------------------------------------
def combine_adjacency(n_times: int, spatial_adj) -> sparse.csr_matrix:
    """index = t * n_ch + ch.  Edges: (t+-1, ch), (t, spatial neighbours of ch) and the
    diagonal; no diagonal-in-time neighbours (MNE connects nodes that differ in
    exactly one dimension)."""
    sp = sparse.coo_matrix(spatial_adj)
    n_ch = sp.shape[0]
    keep = sp.row != sp.col
    srow, scol = sp.row[keep], sp.col[keep]
    t = np.arange(n_times)
    rows = [(t[:, None] * n_ch + srow[None, :]).ravel()]
    cols = [(t[:, None] * n_ch + scol[None, :]).ravel()]
    ch = np.arange(n_ch)
    a = (t[:-1, None] * n_ch + ch[None, :]).ravel()
    rows += [a, a + n_ch, np.arange(n_times * n_ch)]
    cols += [a + n_ch, a, np.arange(n_times * n_ch)]
    r = np.concatenate(rows)
    c = np.concatenate(cols)
    n = n_times * n_ch
    m = sparse.coo_matrix((np.ones(len(r)), (r, c)), shape=(n, n)).tocsr()
    m.data[:] = 1.0
    return m


def add_phase_wraparound(adj, n_times: int, n_ch: int) -> sparse.csr_matrix:
    """(0, ch) <-> (n_times - 1, ch) edges, cbpa.py:949-982."""
    ch = np.arange(n_ch)
    first, last = ch, (n_times - 1) * n_ch + ch
    w = sparse.coo_matrix((np.ones(2 * n_ch), (np.r_[first, last], np.r_[last, first])),
                          shape=adj.shape)
    out = ((adj.astype(bool) + w.tocsr().astype(bool)).astype(bool)).tocsr()
    return out


def delaunay_adjacency(pos: np.ndarray) -> sparse.csr_matrix:
    """Symmetric 0/1 neighbour matrix from the Delaunay triangulation of 2-D sensor
    positions - structurally what mne.channels.find_ch_adjacency builds for a
    montage without a template (cbpa.py:235)."""
    from scipy.spatial import Delaunay
    tri = Delaunay(pos)
    n = len(pos)
    r, c = [], []
    for s in tri.simplices:
        for a in range(3):
            for b in range(3):
                if a != b:
                    r.append(s[a])
                    c.append(s[b])
    m = sparse.coo_matrix((np.ones(len(r)), (r, c)), shape=(n, n)).tocsr()
    m.data[:] = 1.0
    return m


# --------------------------------------------------------------------------
# statistic
# --------------------------------------------------------------------------
def ttest_1samp_no_p(X: np.ndarray) -> np.ndarray:
    """mne.stats.ttest_1samp_no_p with sigma=0: mean / sqrt(var(ddof=1) / n)."""
    with np.errstate(divide="ignore", invalid="ignore"):
        var = np.var(X, axis=0, ddof=1)
        return np.mean(X, axis=0) / np.sqrt(var / X.shape[0])


def to_fixed(t: np.ndarray) -> np.ndarray:
    """int64 fixed-point image of a t-map (NaN -> 0; masked nodes are never NaN)."""
    tc = np.clip(np.nan_to_num(t, nan=0.0, posinf=T_CLAMP, neginf=-T_CLAMP), -T_CLAMP, T_CLAMP)
    return np.rint(tc * FIX_SCALE).astype(np.int64)


def _masks(t: np.ndarray, thr: float, tail: int):
    if tail == 0:
        return [t > thr, t < -thr]
    if tail == 1:
        return [t > thr]
    return [t < thr]


def find_clusters(t: np.ndarray, thr: float, tail: int, adj_coo):
    """Clusters in MNE order: positive mask first, inside one mask by smallest flat
    index.  Returns (list of index arrays, float64 masses, int64 fixed masses)."""
    clusters, fmass, imass = [], [], []
    tf = to_fixed(t)
    n = len(t)
    for x_in in _masks(t, thr, tail):
        if not np.any(x_in):
            continue
        keep = x_in[adj_coo.row] & x_in[adj_coo.col]
        idx = np.flatnonzero(x_in)
        row = np.concatenate((adj_coo.row[keep], idx))
        col = np.concatenate((adj_coo.col[keep], idx))
        g = sparse.coo_matrix((np.ones(len(row)), (row, col)), shape=(n, n))
        _, comp = connected_components(g)
        order = np.argsort(comp[idx], kind="stable")
        lab_sorted = comp[idx][order]
        cuts = np.flatnonzero(np.diff(lab_sorted)) + 1
        for c in np.split(idx[order], cuts):
            clusters.append(c)
            fmass.append(np.sum(t[c]))
            imass.append(int(tf[c].sum()))
    return clusters, np.asarray(fmass, dtype=np.float64), np.asarray(imass, dtype=np.int64)


def labels_from_clusters(clusters, n_tests: int) -> np.ndarray:
    """int32 label map: 0 = not in a cluster, k = k-th cluster (1-based, MNE order)."""
    lab = np.zeros(n_tests, dtype=np.int32)
    for k, c in enumerate(clusters):
        lab[c] = k + 1
    return lab


def max_stat_fixed(imass: np.ndarray, tail: int) -> int:
    """Signed fixed-point cluster mass with the largest magnitude (first wins ties:
    ``np.argmax(np.abs(sums))``), 0 when there is no cluster."""
    if len(imass) == 0:
        return 0
    return int(imass[int(np.argmax(np.abs(imass)))])


def permutation_cluster_1samp_test(X: np.ndarray, signs: np.ndarray, threshold: float,
                                   tail: int, adjacency):
    """X (n_subj, ...) -> dict(t_obs, clusters (bool masks shaped like a sample),
    cluster_pv, H0 (1 + n_perm,), labels, mass_fixed, mass_float, H0_fixed)."""
    sample_shape = X.shape[1:]
    Xf = np.ascontiguousarray(X.reshape(X.shape[0], -1), dtype=np.float64)
    n_tests = Xf.shape[1]
    coo = sparse.coo_matrix(adjacency)
    t_obs = ttest_1samp_no_p(Xf)
    clusters, fmass, imass = find_clusters(t_obs, threshold, tail, coo)
    if tail == 0:
        orig = int(np.max(np.abs(imass))) if len(imass) else 0
    elif tail == 1:
        orig = int(np.max(imass)) if len(imass) else 0
    else:
        orig = int(np.min(imass)) if len(imass) else 0
    h0 = np.zeros(1 + len(signs), dtype=np.int64)
    h0[0] = orig
    for p, s in enumerate(np.asarray(signs)):
        tp = ttest_1samp_no_p(Xf * s[:, None].astype(np.float64))
        _, _, im = find_clusters(tp, threshold, tail, coo)
        h0[1 + p] = max_stat_fixed(im, tail)
    pv = pvalues_from_h0(imass, h0, tail)
    masks = []
    for c in clusters:
        m = np.zeros(n_tests, dtype=bool)
        m[c] = True
        masks.append(m.reshape(sample_shape))
    return dict(t_obs=t_obs.reshape(sample_shape), clusters=masks, cluster_pv=pv,
                H0=h0.astype(np.float64) / FIX_SCALE, H0_fixed=h0,
                labels=labels_from_clusters(clusters, n_tests), mass_fixed=imass,
                mass_float=fmass)


def pvalues_from_h0(stats_fixed: np.ndarray, h0_fixed: np.ndarray, tail: int) -> np.ndarray:
    """mne ``_pval_from_histogram`` on the integer images (exact counts / len(H0))."""
    if tail == -1:
        return np.array([np.mean(h0_fixed <= s) for s in stats_fixed], dtype=np.float64)
    if tail == 1:
        return np.array([np.mean(h0_fixed >= s) for s in stats_fixed], dtype=np.float64)
    return np.array([np.mean(np.abs(h0_fixed) >= abs(int(s))) for s in stats_fixed],
                    dtype=np.float64)
