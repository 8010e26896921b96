"""Import the UNMODIFIED reference modules from /root/reference (build container only).

TEST INFRASTRUCTURE.  Used by ``scripts/make_golden.py`` to generate the golden
vectors under ``tests/golden/`` and, when the reference tree is present, by
``bench.py --impl reference``.  The GPU box has no /root/reference: callers must
handle ``reference_available() == False``.

``signal_features.py:9,11`` imports matplotlib and ``visualizations`` (which
forces a Qt backend at ``visualizations.py:55``); neither is installed here and
neither is touched by the numeric path, so they are replaced by inert stubs.
"""
from __future__ import annotations

import os
import sys
from unittest.mock import MagicMock

REFERENCE_ROOT = os.environ.get("CMC_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "src", "pipeline", "signal_features.py"))


def load_reference():
    """Returns (signal_features, data_surrogation) modules of the reference."""
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    for name in ("matplotlib", "matplotlib.pyplot", "src.pipeline.visualizations"):
        sys.modules.setdefault(name, MagicMock())
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import src.pipeline.signal_features as sf        # noqa: E402
    import src.pipeline.data_surrogation as ds       # noqa: E402
    return sf, ds
