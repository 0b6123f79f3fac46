"""Grid collimator (host side, fp64).

Drop-in for ``co_monitor.geometry.collimator.Collimator``
(reference: co_monitor/geometry/collimator.py:10-125).  A collimator is ``S``
open slits, 1 mm wide on a 2 mm pitch: slit ``k`` on the crystal side is the quad

    A1 + 2k t,  A1 + t + 2k t,  A2 + 2k t,  A2 + 0.001 + t + 2k t       (t = vector_top_bottom)

and the plasma side is the same quad moved by ``vector front-back``
(collimator.py:103-119).  The scalar ``0.001`` really is added to all three
components of the fourth vertex in the reference; it makes the quad slightly
non-planar and is kept because it shapes the aperture cross-section.
"""
from __future__ import annotations

import numpy as np

from .json_reader import read_json_file


class Collimator:
    def __init__(self, element: str, closing_side: str, slits_number: int = 10, plot: bool = False):
        self.slits_number = slits_number
        self.element = element
        self.closing_side = closing_side
        self.loaded_file = read_json_file()
        self.collimator, self.vector_front_back = self.get_coordinates()
        self._init_collim_coord()
        if plot:
            self.visualization()

    def __repr__(self):
        return f'Collimator(element="{self.element}", A={self.A1}, B={self.A2}, C={self.B1})'

    def get_coordinates(self):
        entry = self.loaded_file["collimator"]["element"][f"{self.element}"]
        return entry[f"{self.closing_side}"], np.array(entry["vector front-back"])

    def _init_collim_coord(self):
        self.vector_top_bottom = np.array(self.collimator["vector_top_bottom"])
        for name in ("A1", "A2", "B1", "B2"):
            setattr(self, name, np.array(self.collimator["vertex"][name]))

    def spatial_colimator(self, vertices_coordinates: np.ndarray) -> np.ndarray:
        """Slab of one slit: its quad pushed +/- the front-back vector, shape (8, 3)."""
        return np.concatenate(
            (
                vertices_coordinates + self.vector_front_back,
                vertices_coordinates - self.vector_front_back,
            )
        ).reshape(8, 3)

    def read_colim_coord(self):
        """Return ``(S,8,3)`` both sides stacked, ``(S,4,3)`` crystal side, ``(S,4,3)`` plasma side."""
        t = self.vector_top_bottom
        crys = np.empty((self.slits_number, 4, 3))
        for k in range(self.slits_number):
            pitch = t * 2 * k
            crys[k, 0, :] = self.A1 + pitch
            crys[k, 1, :] = (self.A1 + t) + pitch
            crys[k, 2, :] = self.A2 + pitch
            crys[k, 3, :] = (self.A2 + 0.001 + t) + pitch
        plasma = crys + self.vector_front_back
        return np.concatenate((crys, plasma), axis=1), crys, plasma

    def visualization(self):
        raise NotImplementedError(
            "3-D rendering (pyvista) is outside the accelerated path of co_monitor_b200"
        )
