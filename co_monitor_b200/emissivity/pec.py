"""ADAS photon-emissivity-coefficient tables (host side, fp64).

Drop-in for ``co_monitor.emissivity.pec.PEC`` (reference: co_monitor/emissivity/pec.py:13-187).
File choice by the ``"<element><charge>"`` substring (pec.py:63-80), block choice by
``float(tokens[0]) == wavelength and transition in tokens`` (pec.py:94-120), block length
``ceil(n_ne/8) + ceil(n_te/8) + n_ne*ceil(n_te/8)`` lines (pec.py:134-151).

The reference then resamples the block with ``interp2d(kind="linear")`` onto a 2000 x 2000
``linspace`` grid (pec.py:153-187; a 96 MB array per transition).  A degree-1 tensor spline through
all nodes is plain bilinear interpolation, so the device keeps only the ADAS nodes (<= 24 x 24)
and evaluates the bilinear form at the nearest ``linspace`` node (emissivity.cu); the dense table
is still available here as ``interpolated_pec`` (built lazily, NumPy) for callers that want it.
"""
from __future__ import annotations

from itertools import islice
from pathlib import Path

import numpy as np

from ._paths import stage2_input_root


def bilinear_on_grid(x_nodes, y_nodes, table, x_new, y_new) -> np.ndarray:
    """Degree-1 tensor spline through the nodes, evaluated in FITPACK's operation order (fpbspl + fpbisp), which is
    what ``interp2d(kind="linear")`` ran.  ``table`` is (len(x_nodes), len(y_nodes)); returns (len(x_new), len(y_new))."""

    def cell(nodes, q):
        nodes, q = np.asarray(nodes, float), np.asarray(q, float)
        i = np.clip(np.searchsorted(nodes, q, side="right") - 1, 0, len(nodes) - 2)
        f = 1.0 / (nodes[i + 1] - nodes[i])
        return i, f * (nodes[i + 1] - q), f * (q - nodes[i])

    ix, hx0, hx1 = cell(x_nodes, x_new)
    iy, hy0, hy1 = cell(y_nodes, y_new)
    ixc, iyr = ix[:, None], iy[None, :]
    hx0, hx1, hy0, hy1 = hx0[:, None], hx1[:, None], hy0[None, :], hy1[None, :]
    sp = table[ixc, iyr] * hx0 * hy0
    sp = sp + table[ixc, iyr + 1] * hx0 * hy1
    sp = sp + table[ixc + 1, iyr] * hx1 * hy0
    sp = sp + table[ixc + 1, iyr + 1] * hx1 * hy1
    return sp


class PEC:
    def __init__(self, element, wavelength, ionization_state, transition, interp_step=2000, plot=False,
                 input_root=None):
        self.element = element
        self.wavelength = wavelength
        self.ionization_state = ionization_state
        self.transition = transition
        self.interp_step = interp_step
        self._input_root = input_root
        self.file_path = self._get_file_path()
        self.head_idx, self.ne_nodes_nr, self.te_nodes_nr = self._get_header_info()
        self.pec_data_array, self.ne_nodes, self.te_nodes = self._read_data()
        # the uniform grids the reference looks values up on (pec.py:169-174)
        self.ne_grid = np.linspace(self.ne_nodes[0], self.ne_nodes[-1], self.interp_step, endpoint=True)
        self.te_grid = np.linspace(self.te_nodes[0], self.te_nodes[-1], self.interp_step, endpoint=True)
        self._dense = None

    def _get_file_path(self) -> Path:
        directory = stage2_input_root(self._input_root) / "PEC" / self.element
        tag = str(self.element).lower() + str(self.ionization_state)
        matches = [f for f in directory.glob("*") if f.is_file() and tag in f.name]
        return matches[0]

    def _get_header_info(self):
        with open(self.file_path) as fh:
            self.data = fh.readlines()
        head = None
        for idx, line in enumerate(self.data):
            tokens = line.split()
            try:
                if float(tokens[0]) == self.wavelength and self.transition in tokens:
                    head, head_idx = tokens, idx + 1
                    break
            except (ValueError, IndexError):  # comment section: end of the data blocks
                break
        if head is None:
            raise UnboundLocalError(
                f"no {self.transition} block at {self.wavelength} A in {self.file_path.name}")
        return head_idx, int(head[2]), int(head[3])

    def _read_data(self):
        n_ne, n_te = self.ne_nodes_nr, self.te_nodes_nr
        n_rows = int(np.ceil(n_ne / 8) + np.ceil(n_te / 8) + n_ne * np.ceil(n_te / 8))
        with open(self.file_path, "r") as fh:
            values = [v for line in islice(fh, self.head_idx, self.head_idx + n_rows) for v in line.split()]
        data = np.array(values, dtype=np.float64)
        return data[n_ne + n_te:].reshape((n_ne, n_te)), data[:n_ne], data[n_ne:n_ne + n_te]

    @property
    def interpolated_pec(self) -> np.ndarray:
        """``(interp_step, interp_step, 3)`` array of (ne, te, pec), as pec.py:153-187 builds it."""
        if self._dense is None:
            pec = bilinear_on_grid(self.ne_nodes, self.te_nodes, self.pec_data_array, self.ne_grid, self.te_grid)
            ne, te = np.meshgrid(self.ne_grid, self.te_grid, indexing="ij")
            self._dense = np.stack((ne, te, pec), axis=-1)
        return self._dense
