"""Curved dispersive element / crystal (host side, fp64).

Drop-in for ``co_monitor.geometry.dispersive_element.DispersiveElement``
(reference: co_monitor/geometry/dispersive_element.py:8-173).  The crystal is a
cylinder arc sampled on ``height_step x length_step`` points: height
``linspace(0, 20, h)`` along the cylinder axis, arc ``linspace(0, alpha, l)`` at
the *integer-truncated* radius ``int(round(|A - rc|, 2))`` (dispersive_element.py:47,
64-83), then rotated onto the crystal orientation vector ``B - C`` and spun about
it so that the first sample lines up with vertex ``C`` (dispersive_element.py:137-171).
The NumPy operation order is kept so that ``crystal_points`` is bit-identical to
the reference's.
"""
from __future__ import annotations

import numpy as np

from .json_reader import read_json_file
from .rotation_matrix import rotation_matrix


class DispersiveElement:
    def __init__(self, element, crystal_height_step=10, crystal_length_step=20):
        self.element = element
        self.height_step = crystal_height_step
        self.length_step = crystal_length_step
        self.loaded_file = read_json_file()
        self.disp_elem_coord = self.read_coord_from_file()
        self._init_crys_coordinates()
        self.crystal_points = self._make_curved_crystal()

    def read_coord_from_file(self):
        return self.loaded_file["dispersive element"]["element"][f"{self.element}"]

    def _init_crys_coordinates(self):
        db = self.disp_elem_coord
        self.max_reflectivity = db["max reflectivity"]
        self.crystal_central_point = np.array(db["crystal central point"])
        self.radius_central_point = np.array(db["radius central point"])
        self.AOI = np.array(db["AOI"])
        for name in "ABCD":
            setattr(self, name, np.array(db["vertex"][name]))
        self.radius = self.distance_between_points(self.A, self.radius_central_point)
        self.crystal_orientation_vector = self.B - self.C
        self.alpha = self.angle_between_lines(self.radius_central_point, self.A, self.B)

    @staticmethod
    def distance_between_points(p1: np.ndarray, p2: np.ndarray) -> int:
        """Euclidean distance, rounded to 2 decimals and then *truncated* to int (mm)."""
        return int(np.sqrt(np.sum((p1 - p2) ** 2, axis=0)).round(2))

    def angle_between_lines(self, central_point, p1, p2):
        """Angle p1-central-p2 in degrees."""
        central_point, p1, p2 = np.array(central_point), np.array(p1), np.array(p2)
        v1 = p1 - central_point
        v2 = p2 - central_point
        return np.degrees(
            np.arccos(np.dot(v1, v2) / (np.linalg.norm(v1) * np.linalg.norm(v2)))
        )

    def _make_curved_crystal(self) -> np.ndarray:
        heights = np.linspace(0, 20, self.height_step)
        arc = np.linspace(0, self.alpha, num=self.length_step) / 180 * np.pi
        ring_x = self.radius * np.cos(arc)
        ring_y = self.radius * np.sin(arc)
        # arc index is the slow axis, height the fast one; cylinder axis along x
        pts = np.vstack(
            (
                np.tile(heights, self.length_step),
                np.repeat(ring_x, self.height_step),
                np.repeat(ring_y, self.height_step),
            )
        ).T

        axis_dir = self.crystal_orientation_vector / (
            np.linalg.norm(self.crystal_orientation_vector)
        )
        x_hat = np.array([1, 0, 0])
        if np.allclose(x_hat, axis_dir):
            return pts

        tilt = rotation_matrix(np.arccos(np.dot(axis_dir, x_hat)), np.cross(axis_dir, x_hat))
        pts = pts.dot(tilt)

        shift = np.array(self.radius_central_point - self.crystal_orientation_vector / 2)
        pts += shift
        spin_deg = self.angle_between_lines(self.radius_central_point, self.C, pts[0])
        pts -= shift
        spin = rotation_matrix(np.deg2rad(spin_deg), self.crystal_orientation_vector)
        pts = pts.dot(spin)
        pts += shift
        return pts
