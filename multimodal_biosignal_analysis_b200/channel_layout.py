"""Channel-name constants of the 64-channel EEG cap and the 8x8 HD-EMG grid (data, mirrors
``src/pipeline/channel_layout.py:3-34`` of the reference so channel subsets index the same columns)."""

_ROWS = (
    "Fp1 Fpz Fp2",
    "AF7 AF3 AFz AF4 AF8",
    "F9 F7 F3 F1 Fz F2 F4 F8 F10",
    "FT9 FT7",
    "FC5 FC3 FC1 FCz FC2 FC4 FC6",
    "FT8 FT10",
    "T9 T7",
    "C5 C3 C1 Cz C2 C4 C6",
    "T8 T10",
    "TP9 TP7",
    "CP5 CP3 CP1 CPz CP2 CP4 CP6",
    "TP8 TP10",
    "P9 P7 P3 P1 Pz P2 P4 P8 P10",
    "PO7 POz PO8",
    "O1 O2",
)
EEG_CHANNELS = [name for row in _ROWS for name in row.split()]
assert len(EEG_CHANNELS) == 64

_AREAS = [('Frontal Pole', 'Fp'), ('Anterior Frontal', 'AF'), ('Fronto-Central', 'FC'), ('Frontal', 'F'),
          ('Fronto-Temporal', 'FT'), ('Temporal', 'T'), ('Central', 'C'), ('Temporo-Parietal', 'TP'),
          ('Centro-Parietal', 'CP'), ('Parietal', 'P'), ('Parieto-Occipital', 'PO'), ('Occipital', 'O')]


def _in_area(ch: str, abbr: str) -> bool:
    rest = ch[len(abbr):]
    return ch.startswith(abbr) and (rest.isnumeric() or rest == 'z')


EEG_CHANNELS_BY_AREA = {label: [ch for ch in EEG_CHANNELS if _in_area(ch, abbr)] for label, abbr in _AREAS}
EEG_CHANNEL_IND_DICT = {ch: ind for ind, ch in enumerate(EEG_CHANNELS)}
EMG_CHANNELS = [f"EMG{i:02d}" for i in range(64)]

# 2-D sensor layout of the cap (data; same coordinates as ``EEG_POSITIONS`` of the reference's
# ``src/pipeline/visualizations.py:61-131``, to 4 decimals).  The layout is left / right symmetric: midline and
# left-hemisphere (odd) sites are listed, even sites mirror them.  ``cbpa.run_cbpa`` derives its default spatial
# adjacency from these positions (Delaunay neighbours) when no MNE montage is available.
_MIDLINE_Y = {"Fpz": 0.602, "AFz": 0.42, "Fz": 0.252, "FCz": 0.126, "Cz": 0.0, "CPz": -0.126, "Pz": -0.252,
              "POz": -0.42}
_LEFT_XY = {
    "Fp1": (0.165, 0.56), "AF7": (0.308, 0.49), "AF3": (0.154, 0.448), "F9": (0.506, 0.455), "F7": (0.41, 0.385),
    "F3": (0.22, 0.294), "F1": (0.11, 0.266), "FT9": (0.594, 0.238), "FT7": (0.484, 0.196), "FC5": (0.3685, 0.168),
    "FC3": (0.253, 0.147), "FC1": (0.1293, 0.133), "T9": (0.64, 0.0), "T7": (0.53, 0.0), "C5": (0.4125, 0.0),
    "C3": (0.275, 0.0), "C1": (0.1375, 0.0), "TP9": (0.6, -0.24), "TP7": (0.484, -0.196), "CP5": (0.3685, -0.168),
    "CP3": (0.253, -0.147), "CP1": (0.1293, -0.133), "P9": (0.47, -0.42), "P7": (0.37, -0.355), "P3": (0.22, -0.294),
    "P1": (0.11, -0.266), "PO7": (0.308, -0.49), "O1": (0.165, -0.56),
}


def _positions() -> dict:
    import re
    pos = {ch: (0.0, y) for ch, y in _MIDLINE_Y.items()}
    for ch, (ax, y) in _LEFT_XY.items():
        area, num = re.fullmatch(r"([A-Za-z]+)(\d+)", ch).groups()
        pos[ch] = (-ax, y)
        pos[f"{area}{int(num) + 1}"] = (ax, y)
    return {ch: pos[ch] for ch in EEG_CHANNELS}


EEG_POSITIONS = _positions()
assert len(EEG_POSITIONS) == 64
