"""Geometry database loader (host side).

Mirrors ``co_monitor.geometry.json_reader.read_json_file``
(reference: co_monitor/geometry/json_reader.py:7-26), which resolves
``<pkg>/../input_files/geometry/coordinates.json``.  Here the database ships
inside the package (``co_monitor_b200/input_files``); ``COMONITOR_INPUT_ROOT``
or an explicit *path* overrides it.
"""
from __future__ import annotations

import json
import os
from functools import lru_cache
from pathlib import Path

PACKAGE_INPUT_ROOT = Path(__file__).resolve().parent.parent / "input_files"


def input_root() -> Path:
    """Directory that plays the role of the reference's ``input_files/``."""
    env = os.environ.get("COMONITOR_INPUT_ROOT")
    return Path(env) if env else PACKAGE_INPUT_ROOT


@lru_cache(maxsize=8)
def _load(path: str) -> str:
    with open(path) as fh:
        return fh.read()


def read_json_file(path: str | os.PathLike | None = None) -> dict:
    """Return the diagnostic coordinate database as a dict (a fresh copy each call)."""
    if path is None:
        path = input_root() / "geometry" / "coordinates.json"
    return json.loads(_load(str(path)))
