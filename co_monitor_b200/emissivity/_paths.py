"""Where Stage 2 finds its inputs.

The reference resolves everything against ``Path.cwd() / "input_files"`` (pec.py:63,
fractional_abundance.py:48-53, reader.py:116-122,150-157).  Same here when that directory
exists; otherwise the copy that ships inside the package is used, and ``COMONITOR_INPUT_ROOT``
overrides both.
"""
from __future__ import annotations

import os
from pathlib import Path

from ..geometry.json_reader import PACKAGE_INPUT_ROOT


def stage2_input_root(explicit=None) -> Path:
    if explicit is not None:
        return Path(explicit)
    env = os.environ.get("COMONITOR_INPUT_ROOT")
    if env:
        return Path(env)
    cwd = Path.cwd() / "input_files"
    return cwd if cwd.is_dir() else PACKAGE_INPUT_ROOT
