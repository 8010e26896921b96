"""The ctypes stub printed in INTEGRATION.md (section 2) is executed as written and checked against the oracle."""
import os
import re

import numpy as np
import pytest
from scipy.stats import t as t_dist

from oracle import coherence as oc

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_integration_md_ctypes_stub_runs_and_matches_oracle(cuda_device):
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    blocks = re.findall(r"```python\n(.*?)```", text, flags=re.S)
    stub = next(b for b in blocks if "def gpu_window_msc" in b)
    ns = {}
    cwd = os.getcwd()
    os.chdir(ROOT)                                   # the stub loads the library by its in-tree relative path
    try:
        exec(compile(stub, "INTEGRATION.md", "exec"), ns)
        rng = np.random.default_rng(5)
        N, ne, nm = 256, 4, 6
        eeg = rng.standard_normal((N, ne))
        emg = rng.standard_normal((N, nm)) + 0.5 * eeg[:, :1]
        tapers, _ = oc.dpss_tapers(N, 3)
        t_crit = float(t_dist.ppf(0.975, len(tapers) - 1))
        coh, lo, hi = ns["gpu_window_msc"](eeg, emg, tapers, t_crit)
    finally:
        os.chdir(cwd)
    X = oc.segment_spectra(eeg, np.array([0]), tapers)[0]
    Y = oc.segment_spectra(emg, np.array([0]), tapers)[0]
    m, l, h = oc.jackknife_from_spectra(X, Y, 0.05)
    assert coh.shape == m.shape == (N // 2 + 1, ne, nm)
    assert np.max(np.abs(coh - m)) < 1e-4
    tame = m < 0.999
    assert np.max(np.abs(lo - l)[tame]) < 2e-3 and np.max(np.abs(hi - h)[tame]) < 2e-3
