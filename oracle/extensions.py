"""TEST INFRASTRUCTURE - CPU statements of the optional physics extensions.

These switches are OFF on the parity path; the reference computes none of them, so nothing here is pinned by a
reference output ("parity unpinned", extensions only).  Each function states what the corresponding CUDA code does and
cites the reference lines the extension departs from.

  loglog_pec()         bilinear interpolation of log PEC in (log ne, log Te) through the ADAS nodes, instead of the
                       reference's linear interp2d table read at the nearest node (pec.py:153-187, reader.py:316-341)
                       -> csrc/emissivity.cu::pec_loglog

Only tests/ may import this module.
"""
from __future__ import annotations

import numpy as np


def loglog_pec(x_nodes, y_nodes, table, x, y) -> np.ndarray:
    """``pec_interpolation="loglog"``: bilinear in (log x, log y) of log table through the nodes, arguments clamped to
    the node range."""
    x_nodes, y_nodes, table = (np.asarray(a, dtype=float) for a in (x_nodes, y_nodes, table))
    x = np.clip(np.asarray(x, dtype=float), x_nodes[0], x_nodes[-1])
    y = np.clip(np.asarray(y, dtype=float), y_nodes[0], y_nodes[-1])
    i = np.clip(np.searchsorted(x_nodes, x, side="right") - 1, 0, len(x_nodes) - 2)
    j = np.clip(np.searchsorted(y_nodes, y, side="right") - 1, 0, len(y_nodes) - 2)
    u = (np.log(x) - np.log(x_nodes[i])) / (np.log(x_nodes[i + 1]) - np.log(x_nodes[i]))
    v = (np.log(y) - np.log(y_nodes[j])) / (np.log(y_nodes[j + 1]) - np.log(y_nodes[j]))
    c00, c01, c10, c11 = table[i, j], table[i, j + 1], table[i + 1, j], table[i + 1, j + 1]
    L = (1 - u) * (1 - v) * np.log(c00) + (1 - u) * v * np.log(c01) + u * (1 - v) * np.log(c10) + u * v * np.log(c11)
    return np.exp(L)
