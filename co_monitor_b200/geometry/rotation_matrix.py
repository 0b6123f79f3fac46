"""Euler-Rodrigues rotation (host side, fp64).

Same semantics as ``co_monitor.geometry.rotation_matrix.rotation_matrix``
(reference: co_monitor/geometry/rotation_matrix.py:3-33) and pinned by the
reference's own known-answer tests (tests/unit/geometry/test_rotation_matrix.py):
row-vector convention (``points.dot(R)`` rotates by *theta* about *axis*),
``ValueError`` for an axis that is not 3 long, ``TypeError`` for a non-numeric
angle.  The arithmetic order is kept so that meshes built from it are
bit-identical to the reference's.
"""
import numpy as np


def rotation_matrix(theta, axis) -> np.ndarray:
    unit = np.asarray(axis) / np.sqrt(np.dot(axis, axis))
    half = theta / 2.0  # TypeError for a str angle, as in the reference
    q0 = np.cos(half)
    q1, q2, q3 = -unit * np.sin(half)  # ValueError unless len(axis) == 3
    s00, s11, s22, s33 = q0**2, q1**2, q2**2, q3**2
    p12, p03, p02, p01, p13, p23 = q1 * q2, q0 * q3, q0 * q2, q0 * q1, q1 * q3, q2 * q3
    return np.array(
        [
            [s00 + s11 - s22 - s33, 2 * (p12 + p03), 2 * (p13 - p02)],
            [2 * (p12 - p03), s00 + s22 - s11 - s33, 2 * (p23 + p01)],
            [2 * (p13 + p02), 2 * (p23 - p01), s00 + s33 - s11 - s22],
        ]
    )
