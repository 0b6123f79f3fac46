"""TEST INFRASTRUCTURE - only used by the fixture generators in this directory.

Makes the unmodified reference (``/root/reference``, Python) importable in the
build container: stubs the plot-only third-party modules it imports at module
level (pyvista, matplotlib) and, for Stage 2, restores the removed
``scipy.interpolate.interp2d`` with the regular-grid code path it used to take
(degree-1 ``RectBivariateSpline`` == bilinear).
Nothing here is imported by the product or by the tests: the reference does not
exist on the GPU box, so its outputs travel as committed fixtures (tests/golden).
"""
from __future__ import annotations

import sys
import types

REFERENCE_ROOT = "/root/reference"


def install_stubs() -> None:
    if "pyvista" not in sys.modules:
        pv = types.ModuleType("pyvista")

        class _Dummy:
            def __init__(self, *a, **k):
                pass

            def __getattr__(self, name):
                return lambda *a, **k: None

        pv.PolyData = _Dummy
        pv.Plotter = _Dummy
        core = types.ModuleType("pyvista.core")
        pointset = types.ModuleType("pyvista.core.pointset")
        pointset.PolyData = _Dummy
        core.pointset = pointset
        pv.core = core
        sys.modules["pyvista"] = pv
        sys.modules["pyvista.core"] = core
        sys.modules["pyvista.core.pointset"] = pointset
    if "matplotlib" not in sys.modules:
        mpl = types.ModuleType("matplotlib")
        plt = types.ModuleType("matplotlib.pyplot")
        plt.__getattr__ = lambda name: (lambda *a, **k: None)  # type: ignore[attr-defined]
        mpl.pyplot = plt
        sys.modules["matplotlib"] = mpl
        sys.modules["matplotlib.pyplot"] = plt
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)


def install_interp2d() -> None:
    """scipy >= 1.14 removed interp2d; the reference calls it with kind='linear' on a regular grid."""
    import numpy as np
    from scipy import interpolate

    if getattr(interpolate, "_comonitor_interp2d", False):
        return

    class interp2d:  # noqa: N801 - name dictated by scipy
        def __init__(self, x, y, z, kind="linear"):
            assert kind == "linear"
            self._spl = interpolate.RectBivariateSpline(
                np.asarray(x), np.asarray(y), np.asarray(z).T, kx=1, ky=1, s=0
            )

        def __call__(self, xn, yn):
            return self._spl(np.asarray(xn), np.asarray(yn)).T

    interpolate.interp2d = interp2d
    interpolate._comonitor_interp2d = True
