"""Stage 2 drop-in: ``Emissivity`` with the reference's call signature, attributes and output file.

Replaces ``co_monitor.emissivity.reader.Emissivity`` (reference: co_monitor/emissivity/reader.py:22-418).
File handling and the table classes stay on the host (pandas / NumPy, fp64); the per-voxel work -
nearest profile row (:199-243), nearest fractional-abundance temperature (:245-314), nearest PEC grid
node per transition (:316-341), emissivities (:343-378) and the flux sums (:380-401) - runs in the
sm_100a kernel behind ``com_emissivity_integrate``.

Things kept exactly as the reference does them, because results depend on them:
  * ``impurity_concentration`` is a percentage and is divided by 100 (:132-134);
  * lookups are nearest-neighbour with numpy's first-minimum tie rule (:263-280), pinned by the
    reference's own unit test of ``_find_nearest``;
  * rows with NaN Reff are dropped (:176) and rows with ``Reff >= min(max mapped, max profile)`` too (:240-242);
  * the result frame is ``sort_values("Reff [m]")``-ed (:375) and ``total_emissivity`` is the
    ``f"{x:.2e}"`` *string* (:396);
  * all paths are relative to ``Path.cwd()`` when an ``input_files`` directory is there (:116-122, 150-157);
    otherwise the copy shipped with the package is used.
Deviations: the fractional-abundance file is found under both ``.txt`` and ``.dat``;
``pec_data`` (the 2 x 96 MB dense PEC arrays) is built only when somebody reads it.  A missing
``Reff_coordinates-<s>_mm`` file raises ``FileNotFoundError`` as in the reference unless the caller opts in with
``reff_fallback="upsample"``: the shipped 15 mm grid is then resampled onto the lattice on the device
(``com_reff_resample``), with a ``UserWarning`` - the physics input is then synthetic - and cached per process.
Extension, off by default: ``pec_interpolation="loglog"`` evaluates the PEC by bilinear interpolation in log-log space
through the ADAS nodes at each row's own (ne, Te) instead of the reference's nearest node of a linearly interpolated
2000 x 2000 table (pec.py:153-187); results then differ from the reference's by design.
"""
from __future__ import annotations

import re
from pathlib import Path, PurePath

import numpy as np
import pandas as pd

from ._paths import stage2_input_root
from .engine import cached_line_tables
from .fractional_abundance import FractionalAbundance
from .kinetic_profiles import ExperimentalProfile, TwoGaussSumProfile  # noqa: F401  (re-exported like the reference)
from .pec import PEC  # noqa: F401


#: Reff tables synthesised by ``reff_fallback="upsample"``, keyed by (spacing, input root)
_UPSAMPLED_REFF: dict = {}


class Emissivity:
    def __init__(
        self,
        element,
        ion_state,
        wavelength,
        impurity_concentration,
        transitions,
        reff_magnetic_config,
        kinetic_profiles,
        time="Test",
        plot=False,
        *,
        input_root=None,
        output_dir=None,
        observed_plasma_volume=None,
        device=None,
        save=True,
        verbose=True,
        pec_interpolation="reference",
        reff_fallback=None,
    ):
        self._say = print if verbose else (lambda *a, **k: None)
        self._input_root = stage2_input_root(input_root)
        self._output_dir = output_dir
        self._device = device
        self._pec_dense = None
        if reff_fallback not in (None, "upsample"):
            raise ValueError("reff_fallback must be None or 'upsample'")
        self._reff_fallback = reff_fallback
        self.element = element
        self.ionization_state = ion_state
        self.wavelength = wavelength
        self.transitions = transitions
        self.impurity_concentration = self._convert_imp_frac(impurity_concentration)
        self.reff_magnetic_config = reff_magnetic_config
        self.kinetic_profiles = kinetic_profiles
        self.time = time

        self.reff_coordinates = self._get_Reff()
        self.observed_plasma_volume = (
            observed_plasma_volume if observed_plasma_volume is not None else self.load_observed_plasma()
        )
        self.plas_coords_with_rad_frac = self.get_plas_coords_with_rad_frac(plot=False)
        self.df_sel_frac_ab = self.get_frac_ab()

        self._tables = cached_line_tables(element, ion_state, wavelength, tuple(transitions), str(self._input_root),
                                          pec_interpolation)
        self._run_device_pass()
        self.total_emissivity = self.calculate_total_emissivity()
        if save:
            self.savefile()
        if plot:
            self.plot()

    # ---- inputs (host) ------------------------------------------------------------------------------------
    def _convert_imp_frac(self, impurity_fraction):
        return impurity_fraction / 100

    def _get_Reff(self) -> pd.DataFrame:
        path = self._input_root / "geometry" / "reff" / f"{self.reff_magnetic_config}.dat"
        if not path.exists():
            m = re.fullmatch(r"Reff_coordinates-(\d+)_mm", str(self.reff_magnetic_config))
            if not m or self._reff_fallback != "upsample":
                raise FileNotFoundError(path)  # what the reference does (reader.py:116-128)
            import warnings

            from ..geometry.reff import upsample_reff

            warnings.warn(f"{path.name} not found: Reff is SYNTHETIC - the shipped 15 mm grid resampled onto the "
                          f"{m.group(1)} mm lattice (reff_fallback='upsample'; no VMEC service here)", UserWarning, stacklevel=3)
            key = (int(m.group(1)), str(self._input_root))
            if key not in _UPSAMPLED_REFF:
                _UPSAMPLED_REFF[key] = upsample_reff(15, key[0], self._input_root)
            df = _UPSAMPLED_REFF[key].copy()
        else:
            df = pd.read_csv(path, sep=";")
        df.columns = ["idx_plasma", "x", "y", "z", "Reff [m]"]
        return df.astype({"idx_plasma": int, "x": float, "y": float, "z": float, "Reff [m]": float})

    def load_observed_plasma(self) -> pd.DataFrame:
        path = (self._input_root / "geometry" / "observed_plasma_volume" / f"{self.element}"
                / f"{self.element}_plasma_coordinates-10_mm_spacing-height_30-length_20-slit_100.dat")
        return pd.read_csv(path, sep=";")

    def get_plas_coords_with_rad_frac(self, plot=False) -> pd.DataFrame:
        df = self.observed_plasma_volume
        idx_list = df["idx_sel_plas_points"].tolist()
        df["Reff [m]"] = self.reff_coordinates.iloc[idx_list]["Reff [m]"].tolist()
        return df.dropna()

    def get_frac_ab(self) -> pd.DataFrame:
        frac_ab = FractionalAbundance(self.element, self.ionization_state,
                                      input_root=self._input_root).df_interpolated_frac_ab
        cols = np.asarray([frac_ab.iloc[:, 0], frac_ab[self.ionization_state], frac_ab.iloc[:, -1]]).T
        return pd.DataFrame(cols, columns=["T_e [eV]", f"{self.ionization_state}", "Fully stripped"])

    @staticmethod
    def _find_nearest(array, value) -> int:
        """Index of the element of *array* closest to *value*; the first one wins a tie (reader.py:263-280)."""
        return (np.abs(np.array(array) - value)).argmin()

    # ---- device pass -----------------------------------------------------------------------------------------
    def _run_device_pass(self):
        base = self.plas_coords_with_rad_frac.reset_index(drop=True)
        prof = self.kinetic_profiles[["Reff [m]", "T_e [eV]", "n_e [m-3]"]].to_numpy(dtype=float)
        max_mapped = base["Reff [m]"].max()
        self._say(f"\nMax reff from mapped vmec: {max_mapped}")
        out = self._tables.integrate(prof, self.impurity_concentration, base["Reff [m]"].to_numpy(dtype=float),
                                     base["total_intensity_fraction"].to_numpy(dtype=float), per_voxel=True,
                                     device=self._device)
        self.reff_boundary = out["boundary"]
        self._say(f"\nMax reff from profiles: {self.reff_boundary}")
        self._flux = out["flux"]
        keep = out["valid"]
        cols = out["columns"][keep]
        df = base[keep].reset_index(drop=True)
        nt = len(self.transitions)
        df["n_e [m-3]"] = cols[:, 0]
        df["T_e [eV]"] = cols[:, 1]
        self.ne_te_profiles = df.copy()
        df[f"{self.ionization_state}"] = cols[:, 2]
        df["Fully stripped"] = cols[:, 3]
        for t, tr in enumerate(self.transitions):
            df[f"pec_{tr}"] = cols[:, 4 + t]
        for t, tr in enumerate(self.transitions):
            name = "Emissivity_RECOM" if "RECOM" in tr else f"Emissivity_{tr}"
            df[name] = cols[:, 4 + nt + t]
        df["Emissivity_TOTAL"] = cols[:, 4 + 2 * nt]
        df.sort_values("Reff [m]", inplace=True)  # reader.py:375
        self.df_prof_frac_ab_pec = df
        self.df_prof_frac_ab_pec_emissivity = df

    @property
    def pec_data(self) -> np.ndarray:
        """``(T, 2000, 2000, 3)`` dense PEC arrays of the reference (reader.py:83-97); built on first access."""
        if self._pec_dense is None:
            self._pec_dense = np.array([p.interpolated_pec for p in self._tables.pecs])
        return self._pec_dense

    def calculate_total_emissivity(self) -> str:
        for tr, value in zip(self.transitions, self._flux):
            self._say(f"{tr} is: {value:.2e}")
        total = f"{self._flux[len(self.transitions)]:.2e}"
        self._say(f"TOTAL is: {total}")
        return total

    @property
    def flux(self) -> dict:
        """Unrounded per-transition photon flux (+ "TOTAL")."""
        out = {tr: float(v) for tr, v in zip(self.transitions, self._flux)}
        out["TOTAL"] = float(self._flux[len(self.transitions)])
        return out

    def savefile(self):
        directory = Path(self._output_dir) if self._output_dir is not None else (
            Path.cwd() / "results" / "numerical_results")
        directory.mkdir(parents=True, exist_ok=True)
        path = PurePath(directory, f"Emissivity ({self.element}_{self.ionization_state}_{self.time}).dat")
        np.savetxt(path, self.df_prof_frac_ab_pec,
                   header=f"{self.df_prof_frac_ab_pec.columns} \n Total emissivity: {self.total_emissivity}")
        return path

    def plot(self, savefig=False):
        import warnings

        warnings.warn("co_monitor_b200: Emissivity.plot ignored (matplotlib rendering is outside the accelerated path)",
                      stacklevel=2)
