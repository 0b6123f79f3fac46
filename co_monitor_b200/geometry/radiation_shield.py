"""ECRH stray-radiation shields (host side, fp64).

Same data and constructor as ``co_monitor.geometry.radiation_shield.RadiationShield`` (reference:
co_monitor/geometry/radiation_shield.py:9-101): per shield a centre, a radius and an orientation vector from the
geometry database, turned into the vertices of a thin cylinder (1 mm long, 360 points per rim).  The reference only
draws these objects (visualization.py:297-308, 409-417).  The accelerated path can also use them as circular stops -
an extension that is off by default (``Simulation(..., ecrh_shields=True)`` / ``make_channel(shields=...)``): every
ray must then also pass through the rim polygon of each shield of its chamber, tested where the ray's line meets the
mid-plane of the cylinder.  The mesh / PolyData members of the reference are plot-only and are not reproduced.
"""
from __future__ import annotations

import numpy as np

from .json_reader import read_json_file
from .rotation_matrix import rotation_matrix

CHAMBERS = ("upper chamber", "bottom chamber")
SHIELDS = ("1st shield", "2nd shield")


class RadiationShield:
    def __init__(self, chamber_position: str, selected_shield: str):
        self.chamber_position = chamber_position
        self.selected_shield = selected_shield
        self.loaded_file = read_json_file()
        self.shields_coordinates = self.get_coordinates()
        self.radius_central_point = np.array(self.shields_coordinates["central point"], dtype=float)
        self.radius = float(self.shields_coordinates["radius"])
        self.orientation_vector = np.array(self.shields_coordinates["orientation vector"], dtype=float)
        self.vertices_coordinates = self.make_cyllinder()

    def get_coordinates(self) -> dict:
        return self.loaded_file["ECRH shield"][self.chamber_position][self.selected_shield]

    def _frame(self) -> np.ndarray:
        """Rows = images of the cylinder's own (axis, x, y) directions (radiation_shield.py:87-99)."""
        along = self.orientation_vector / np.linalg.norm(self.orientation_vector)
        own_axis = np.array([1, 0, 0])
        if np.allclose(own_axis, along):
            return np.eye(3)
        return rotation_matrix(np.arccos(np.dot(along, own_axis)), np.cross(along, own_axis))

    def make_cyllinder(self, length=1, nlength=2, alpha=360, nalpha=360) -> np.ndarray:
        """Rim vertices of the cylinder, reference order: per rim angle the *nlength* axial stations."""
        stations = np.linspace(0, length, nlength)
        full_turn = int(alpha) == 360
        angles = np.linspace(0, alpha, num=nalpha, endpoint=not full_turn) / 180 * np.pi
        own = np.vstack((np.tile(stations, nalpha), np.repeat(self.radius * np.cos(angles), nlength),
                         np.repeat(self.radius * np.sin(angles), nlength))).T
        frame = self._frame()
        if frame is not None and not np.array_equal(frame, np.eye(3)):
            return own.dot(frame) + self.radius_central_point
        return own  # the reference returns the unshifted points in this (never occurring) case

    def as_stop(self, length=1, nalpha=360) -> dict:
        """The shield as a circular stop for ``ChannelScene(shields=[...])``: mid-plane point, unit axis, in-plane basis
        with rim vertex 0 along +u, circumradius and number of rim sides."""
        frame = self._frame()
        return dict(centre=self.radius_central_point + 0.5 * length * frame[0], axis=frame[0], u=frame[1], v=frame[2],
                    radius=self.radius, n_sides=int(nalpha))


def ecrh_shields(element: str) -> list[dict]:
    """The two shields of the chamber the channel *element* sits in, as stops.  The database does not name the chamber
    of a channel: it is the one whose shields lie nearest the channel's crystal (C and O above, B and N below)."""
    crystal = np.array(read_json_file()["dispersive element"]["element"][element]["crystal central point"], dtype=float)
    per_chamber = {ch: [RadiationShield(ch, sh) for sh in SHIELDS] for ch in CHAMBERS}
    chamber = min(CHAMBERS, key=lambda ch: min(np.linalg.norm(s.radius_central_point - crystal) for s in per_chamber[ch]))
    return [s.as_stop() for s in per_chamber[chamber]]
