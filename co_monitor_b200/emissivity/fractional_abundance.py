"""Fractional-abundance tables (host side, fp64).

Drop-in for ``co_monitor.emissivity.fractional_abundance.FractionalAbundance`` (reference:
co_monitor/emissivity/fractional_abundance.py:12-91): the 800-row log-spaced table (5 ... 20000 eV,
per cent) is linearly interpolated onto ``arange(T[0], T[-1], step)`` and divided by 100.

The reference opens ``fractional_abundance_<el>.txt`` while the shipped files are ``.dat``
(FileNotFoundError at HEAD); both names are accepted here.
"""
from __future__ import annotations

import numpy as np
import pandas as pd

from ._paths import stage2_input_root


def interp1d_linear(x, y, x_new):
    """scipy.interpolate.interp1d(kind='linear') for sorted x, same arithmetic (slope * dx + y_lo)."""
    x, y, x_new = np.asarray(x, float), np.asarray(y, float), np.asarray(x_new, float)
    hi = np.clip(np.searchsorted(x, x_new), 1, len(x) - 1)
    lo = hi - 1
    slope = (y[hi] - y[lo]) / (x[hi] - x[lo])
    return slope * (x_new - x[lo]) + y[lo]


class FractionalAbundance:
    def __init__(self, element, ionization_state, interpolation_step=1, plot=False, input_root=None):
        self.element = element
        self.ionization_state = ionization_state
        self.interpolation_step = interpolation_step
        self.plot = plot
        self._input_root = input_root
        self.loaded_file_df = self._read_file()
        self.df_interpolated_frac_ab = self._interpolate()

    def _read_file(self) -> pd.DataFrame:
        directory = stage2_input_root(self._input_root) / "fractional_abundance"
        for suffix in (".txt", ".dat"):
            path = directory / f"fractional_abundance_{self.element}{suffix}"
            if path.exists():
                return pd.read_csv(path, delimiter=" ")
        raise FileNotFoundError(directory / f"fractional_abundance_{self.element}.txt")

    def _interpolate(self) -> pd.DataFrame:
        df = self.loaded_file_df
        temp = df["T"].to_numpy(dtype=float)
        Te = np.arange(temp[0], temp[-1], self.interpolation_step)
        out = {"T": Te}
        for name in df.columns[1:]:
            out[name] = interp1d_linear(temp, df[name].to_numpy(dtype=float), Te) / 100
        return pd.DataFrame(out, columns=list(df.columns))
