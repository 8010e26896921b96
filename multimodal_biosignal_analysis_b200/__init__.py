"""B200-native cortico-muscular coherence hot path (drop-in for the reference's
``src/pipeline/signal_features.py``, ``data_surrogation.py`` and ``cbpa.py`` call surface).

Importing the package does not need a GPU; calling a kernel does, and there is no CPU
fallback - see ``_lib.py``.
"""
__version__ = "0.1.0"
