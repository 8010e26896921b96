"""On-disk naming of stored spectrograms - the wire format between feature extraction and CBPA
(``src/utils/file_management.py:9-125`` of the reference): ``"%Y-%m-%d %H_%M_%S <title><suffix>"``,
looked up again by newest timestamp among the files whose name contains every keyword."""
from __future__ import annotations

import os
from datetime import datetime
from pathlib import Path
from typing import Literal

_STAMP = '%Y-%m-%d %H_%M_%S'


def file_title(title: str, dtype_suffix=".svg", short=False) -> str:
    stamp = datetime.now().strftime('%Y%m%d' if short else _STAMP)
    return f"{stamp} {title}{dtype_suffix}"


def most_recent_file(directory, suffix_to_consider: str | None = None,
                     file_title_keywords: list[str] | str | None = None,
                     search_by: Literal["file-title", "meta-data"] = "file-title",
                     return_type: Literal["dict", "latest_file_path"] = "latest_file_path"):
    if search_by not in ("file-title", "meta-data"):
        raise ValueError(f"search_by must be 'file-title' or 'meta-data', got {search_by}")
    directory = Path(directory)
    if not directory.is_dir():
        raise ValueError(f"Provided path {directory} is not a directory!")
    if isinstance(file_title_keywords, str):
        file_title_keywords = [file_title_keywords]
    found = []
    for entry in os.scandir(directory):
        name = entry.name
        if not entry.is_file():
            continue
        if suffix_to_consider is not None:
            if not name.endswith(suffix_to_consider):
                continue
        elif '.DS_Store' in name:
            continue
        if file_title_keywords is not None and not all(k in name for k in file_title_keywords):
            continue
        if search_by == "file-title":
            try:
                when = datetime.strptime(name[:19], _STAMP)
            except ValueError:
                continue            # not one of ours
        else:
            when = entry.stat().st_mtime
        found.append((when, directory / name))
    if not found:
        raise ValueError("Provided directory doesn't contain files matching the provided criteria!")
    found.sort(key=lambda p: p[0], reverse=True)
    if return_type == "latest_file_path":
        return found[0][1]
    return {"files": [f for _, f in found], "dates": [d for d, _ in found]}


def assert_dir(dir_path) -> None:
    Path(dir_path).mkdir(parents=True, exist_ok=True)
