"""Detector aperture (host side, fp64).

Drop-in for ``co_monitor.geometry.detector.Detector`` (reference:
co_monitor/geometry/detector.py:8-112): four corner vertices and a ``(1, 3)``
orientation vector per channel.
"""
import numpy as np

from .json_reader import read_json_file


class Detector:
    def __init__(self, element: str, plot: bool = False):
        self.element = element
        self.loaded_file = read_json_file()
        self.vertices = self.get_vertices(self.element)
        self.orientation_vector = self.get_orientation_vector(self.element)
        self.spatial_det_coordinates = self.create_thick_det(self.vertices)
        if plot:
            self.plotter()

    def get_vertices(self, element: str) -> np.ndarray:
        corners = self.loaded_file["detector"]["element"][element]["vertex"]
        return np.vstack([corners[name] for name in corners])

    def get_orientation_vector(self, element: str) -> np.ndarray:
        vec = np.vstack(
            [self.loaded_file["detector"]["element"][element]["orientation vector"]]
        )
        assert vec.shape == (1, 3), "Orientation vector should consist of 1 point."
        return vec

    def create_thick_det(self, vertices_coordinates: np.ndarray) -> np.ndarray:
        slab = np.concatenate(
            (
                vertices_coordinates + self.orientation_vector / 2,
                vertices_coordinates - self.orientation_vector / 2,
            )
        )
        assert slab.shape == (8, 3), "Wrong number of vertices!"
        return slab

    def plotter(self):
        raise NotImplementedError(
            "3-D rendering (pyvista) is outside the accelerated path of co_monitor_b200"
        )
