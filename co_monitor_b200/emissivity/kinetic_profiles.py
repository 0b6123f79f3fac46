"""Kinetic (n_e, T_e) profiles over Reff (host side, fp64).

Drop-in for ``co_monitor.emissivity.kinetic_profiles`` (reference:
co_monitor/emissivity/kinetic_profiles.py:63-124 TwoGaussSumProfile, :279-458 ExperimentalProfile,
:133-270 TestExperimentalProfile).  All three expose ``profiles_df`` with the columns
``Reff [m]``, ``T_e [eV]``, ``n_e [m-3]`` (the density column carries cm^-3 values).
Plotting / save helpers of the reference are not reproduced.
"""
from __future__ import annotations

import numpy as np
import pandas as pd

from ._paths import stage2_input_root
from .fractional_abundance import interp1d_linear

COLUMNS = ["Reff [m]", "T_e [eV]", "n_e [m-3]"]


class Profile:
    def plot(self, *a, **k):
        raise NotImplementedError("plotting (matplotlib) is outside the accelerated path of co_monitor_b200")


class TwoGaussSumProfile(Profile):
    def __init__(self, ne_coeff, te_coeff, max_Reff=0.539, plot=False):
        self.max_Reff = max_Reff
        self.Reff = np.arange(0, self.max_Reff, 0.001)
        self.ne_coeff = ne_coeff
        self.te_coeff = te_coeff
        self.profiles_df = self.create_df()

    def profile_equation(self, two_gauss_coefficients) -> np.ndarray:
        A1, x1, w1, A2, x2, w2 = two_gauss_coefficients
        return A1 * np.exp(-((self.Reff - x1) ** 2) / (2 * w1**2)) + A2 * np.exp(
            -((self.Reff - x2) ** 2) / (2 * w2**2)
        )

    def create_df(self) -> pd.DataFrame:
        ne = self.profile_equation(self.ne_coeff)
        te = self.profile_equation(self.te_coeff)
        return pd.DataFrame({COLUMNS[0]: self.Reff, COLUMNS[1]: te, COLUMNS[2]: ne}, columns=COLUMNS)


def _resample(reff, ne, te, max_Reff, interp_step) -> pd.DataFrame:
    """Shared tail of both experimental classes (kinetic_profiles.py:246-268, 436-456)."""
    reff, ne, te = (np.asarray(a, dtype=float) for a in (reff, ne, te))
    reff_max = reff[~(reff > max_Reff)].max()
    grid = np.linspace(0, reff_max, interp_step, endpoint=True)
    te_i = interp1d_linear(reff, te, grid)
    ne_i = interp1d_linear(reff, ne, grid) / 1e6
    data = np.round(np.vstack((grid, te_i, ne_i)), 3).T
    return pd.DataFrame(data, columns=COLUMNS)


class ExperimentalProfile(Profile):
    """Fitted profiles from a W7-X profile report (``kinetic_profiles/experimental/<name>.txt``)."""

    def __init__(self, file_name, max_Reff=0.539, interp_step=10000, plot=False, input_root=None):
        self.file_name = file_name
        self.max_Reff = max_Reff
        self.interp_step = interp_step
        self.file_path = stage2_input_root(input_root) / "kinetic_profiles" / "experimental" / f"{file_name}.txt"
        self.te_idx_range, self.ne_idx_range = self._get_index_ranges()
        self.te_section, self.ne_section = self._get_data_from_file()
        self.profiles_df = self._interpolate()

    def _get_index_ranges(self):
        te_index = ne_index = -1
        with open(self.file_path, "r") as fh:
            for idx, line in enumerate(fh):  # the LAST matching headers win, as in the reference
                if "T_e (keV)" in line:
                    te_index = idx - 1
                if "n_e (10^{19}m^{-3})" in line:
                    ne_index = idx - 1
        if te_index == -1 or ne_index == -1:
            print("No electron temperature (t_e) / density (n_e) profile available!")
            return [], []
        n_lines = abs(ne_index - te_index)
        return [te_index, ne_index], [ne_index, ne_index + n_lines]

    def _get_data_from_file(self):
        te_section, ne_section = [], []
        with open(self.file_path, "r") as fh:
            for idx, line in enumerate(fh):
                if line.startswith("#"):
                    continue
                if self.te_idx_range[0] <= idx <= self.te_idx_range[1]:
                    te_section.append(line.split())
                if self.ne_idx_range[0] <= idx <= self.ne_idx_range[1]:
                    ne_section.append(line.split())
        return te_section, ne_section

    def _merge_to_df(self) -> pd.DataFrame:
        te_arr = np.array(self.te_section)[:, [0, 3]]
        ne_arr = np.array(self.ne_section)[:, 3].reshape(-1, 1)
        df = pd.DataFrame(np.concatenate((te_arr, ne_arr), axis=1).astype(float), columns=COLUMNS)
        df["T_e [eV]"] = df["T_e [eV]"] * 1e3
        df["n_e [m-3]"] = df["n_e [m-3]"] * 1e19
        return df

    def _interpolate(self) -> pd.DataFrame:
        df = self._merge_to_df()
        return _resample(df["Reff [m]"], df["n_e [m-3]"], df["T_e [eV]"], self.max_Reff, self.interp_step)


class TestExperimentalProfile(Profile):
    """Time-resolved fits (``kinetic_profiles/experimental_new/<name>-{n_e,T_e}.csv``, tab separated)."""

    __test__ = False  # not a pytest class

    def __init__(self, file_name, time, max_Reff=0.539, interp_step=10000, plot=False, input_root=None):
        self.file_name = file_name
        self.time = time
        self.max_Reff = max_Reff
        self.interp_step = interp_step
        root = stage2_input_root(input_root) / "kinetic_profiles" / "experimental_new"
        self.file_path = (root / f"{file_name}-n_e.csv", root / f"{file_name}-T_e.csv")
        self.profiles_df = self._interpolate()

    def _get_data_from_file(self) -> pd.DataFrame:
        ne = pd.read_csv(self.file_path[0], sep="\t")
        te = pd.read_csv(self.file_path[1], sep="\t")
        df = pd.DataFrame({"Reff [m]": te["rho"] * 0.539, "n_e [m-3]": ne[self.time], "T_e [eV]": te[self.time]})
        return df.loc[df["Reff [m]"] < 0.539]

    def _interpolate(self) -> pd.DataFrame:
        df = self._get_data_from_file()
        return _resample(df["Reff [m]"], df["n_e [m-3]"], df["T_e [eV]"], self.max_Reff, self.interp_step)
