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
