#!/usr/bin/env python
"""Pin the CBPA oracle / kernels to MNE-Python the day ``import mne`` works (it does not in the build container:
the reference lists mne unpinned in environment.yml:11 and there is no network).

    python scripts/make_mne_golden.py            # writes tests/golden/mne_cbpa.npz, then run pytest

What is recorded, per case (the call the reference makes at src/pipeline/cbpa.py:1027-1042):
  * inputs: X (n_subj, n_times, n_ch), the spatial adjacency, threshold, tail, n_permutations, seed;
  * ``mne.stats.combine_adjacency(n_times, spatial)`` as COO (pins cbpa.combine_adjacency, reference :237);
  * t_obs, cluster masks (out_type="mask"), cluster p-values and H0 of
    ``mne.stats.spatio_temporal_cluster_1samp_test``;
  * the sign-flip table MNE actually used, captured by wrapping ``mne.stats.cluster_level._do_1samp_permutations``
    (its ``orders`` argument; signs = 2 * order - 1 as MNE applies them), so that the host-supplied sign table of this
    repo reproduces MNE's permutations exactly.  If the private hook moved in the installed MNE version, the table is
    left out and the tests compare only the observed clustering and the distribution of H0.
tests/test_mne_golden.py skips while the file is absent."""
from __future__ import annotations

import os
import sys

import numpy as np
from scipy import sparse
from scipy.stats import t as t_dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "tests", "golden", "mne_cbpa.npz")


def main():
    try:
        import mne
        from mne.stats import cluster_level, combine_adjacency, spatio_temporal_cluster_1samp_test
    except Exception as exc:                                     # noqa: BLE001
        print(f"mne is not importable here ({type(exc).__name__}: {exc}); nothing written")
        return 1
    from multimodal_biosignal_analysis_b200 import synthetic as syn
    from oracle import cbpa as ocb
    cases = {
        "cfg4_small": dict(shape=(12, 20, 16), tail=0, n_perm=256, seed=42, wrap=False),
        "production": dict(shape=(13, 36, 11), tail=0, n_perm=512, seed=42, wrap=True),
        "one_tailed": dict(shape=(10, 15, 16), tail=1, n_perm=128, seed=7, wrap=False),
    }
    out = {"mne_version": np.array(mne.__version__)}
    for name, c in cases.items():
        n_subj, n_times, n_ch = c["shape"]
        X = syn.make_cbpa_contrast(n_subj, n_times, n_ch, seed=11)
        spatial = ocb.delaunay_adjacency(syn.sensor_positions(64)[:n_ch])
        adj = combine_adjacency(n_times, sparse.csr_matrix(spatial))
        if c["wrap"]:                                            # reference cbpa.py:949-982
            ch = np.arange(n_ch)
            first, last = ch, (n_times - 1) * n_ch + ch
            wrap = sparse.coo_matrix((np.ones(2 * n_ch, bool), (np.r_[first, last], np.r_[last, first])), shape=adj.shape)
            adj = (adj.astype(bool) + wrap.tocsr()).astype(bool)
        q = 0.025 if c["tail"] == 0 else 0.05
        thr = float(t_dist.ppf(1 - q, n_subj - 1))
        captured = {}
        hook = getattr(cluster_level, "_do_1samp_permutations", None)
        if hook is not None:
            def spy(*args, _orig=hook, **kwargs):
                import inspect
                bound = inspect.signature(_orig).bind(*args, **kwargs)
                captured.setdefault("orders", []).append(np.array(bound.arguments["orders"]))
                return _orig(*args, **kwargs)
            cluster_level._do_1samp_permutations = spy
        try:
            t_obs, clusters, pv, H0 = spatio_temporal_cluster_1samp_test(
                X, n_permutations=c["n_perm"], threshold=thr, tail=c["tail"], adjacency=adj, n_jobs=1,
                seed=np.random.default_rng(c["seed"]), out_type="mask", verbose=False)
        finally:
            if hook is not None:
                cluster_level._do_1samp_permutations = hook
        coo = sparse.coo_matrix(adj)
        out.update({f"{name}_X": X, f"{name}_adj_row": coo.row, f"{name}_adj_col": coo.col,
                    f"{name}_spatial": np.asarray(sparse.csr_matrix(spatial).todense()),
                    f"{name}_thr": thr, f"{name}_tail": c["tail"], f"{name}_t_obs": t_obs,
                    f"{name}_clusters": np.stack(clusters) if len(clusters) else np.zeros((0, n_times, n_ch), bool),
                    f"{name}_pv": np.asarray(pv), f"{name}_H0": np.asarray(H0), f"{name}_wrap": c["wrap"]})
        if captured.get("orders"):
            orders = np.concatenate(captured["orders"], axis=0)
            out[f"{name}_signs"] = (2 * orders.astype(np.int8) - 1).astype(np.int8)
        print(f"{name}: {len(clusters)} clusters, H0 {np.asarray(H0).shape}, sign table "
              f"{'captured' if f'{name}_signs' in out else 'NOT captured'}")
    np.savez_compressed(OUT, **out)
    print("written", OUT)
    return 0


if __name__ == "__main__":
    sys.exit(main())
