"""W7-X port aperture (host side, fp64).

Drop-in for ``co_monitor.geometry.port.Port`` (reference:
co_monitor/geometry/port.py:8-72): the 14 rim vertices and the orientation
vector from the geometry database.  The hull/PolyData members of the reference
are plot-only and are not reproduced.
"""
import numpy as np

from .json_reader import read_json_file


class Port:
    def __init__(self, plot: bool = False):
        self.loaded_file = read_json_file()
        self.vertices_coordinates = self.get_vertices()
        self.orientation_vector = self.get_orientation_vector()
        self.spatial_port = self.calculate_port_thickness()
        if plot:
            self.plotter()

    def get_vertices(self) -> np.ndarray:
        rim = self.loaded_file["port"]["vertex"]
        return np.vstack([rim[name] for name in rim])

    def get_orientation_vector(self) -> np.ndarray:
        return np.array(self.loaded_file["port"]["orientation vector"])

    def calculate_port_thickness(self) -> np.ndarray:
        half = np.array(self.orientation_vector) / 2
        return np.concatenate(
            (self.vertices_coordinates + half, self.vertices_coordinates - half)
        )

    def plotter(self):
        raise NotImplementedError(
            "3-D rendering (pyvista) is outside the accelerated path of co_monitor_b200"
        )
