"""TEST INFRASTRUCTURE - CPU oracle for Stage 2 (line-emissivity integration).

NumPy/SciPy restatement of ``co_monitor.emissivity.reader.Emissivity`` and of the table classes it
uses, with pandas reduced to array operations; every lookup is the reference's literal
``argmin(|table - value|)`` (first minimum wins, reader.py:263-280):

  pec_table()          pec.py:54-187        block pick, 2000 x 2000 linear resampling (FITPACK degree-1 spline:
                                            interp2d was removed from SciPy, RectBivariateSpline(kx=ky=1, s=0) is
                                            the code path it took on a regular grid)
  fa_table()           fractional_abundance.py:39-91   interp1d onto arange(T0, T_last, 1), / 100
  two_gauss_profile()  kinetic_profiles.py:63-124
  emissivity()         reader.py:99-130,136-197 (Reff join, NaN drop), :199-243 (nearest profile row, boundary cut),
                       :245-314 (nearest FA temperature), :316-341 (nearest ne / Te grid index -> PEC),
                       :343-378 (emissivities), :380-401 (flux sums)

PARITY PIN: tests/golden/stage2_reference.npz holds the outputs of the UNMODIFIED reference class run in the
build container through oracle/_refshim.py (generator: oracle/make_golden_stage2.py) for B/C/N/O with the
two-Gaussian profile, C with an experimental report, O with a time-resolved fit and three slices of the
temperature sweep; tests/test_oracle_stage2.py checks this restatement against all of them (every per-voxel
column to <= 1e-13 relative, identical row sets).  The reference's own unit test of ``_find_nearest``
(tests/unit/emissivity/test_emissivity.py) is restated there too.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
"""
from __future__ import annotations

import os
from itertools import islice
from types import SimpleNamespace

import numpy as np

DEFAULT_INPUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "co_monitor_b200", "input_files")
COLUMNS = ["idx_sel_plas_points", "plasma_x", "plasma_y", "plasma_z", "total_intensity_fraction", "Reff [m]",
           "n_e [m-3]", "T_e [eV]", None, "Fully stripped", "pec_EXCIT", "pec_RECOM", "Emissivity_EXCIT",
           "Emissivity_RECOM", "Emissivity_TOTAL"]


def find_nearest(array, value) -> int:
    """reader.py:263-280."""
    return int((np.abs(np.array(array) - value)).argmin())


def nearest_rows(table: np.ndarray, values: np.ndarray, chunk: int = 512) -> np.ndarray:
    """find_nearest for many values (same arithmetic, evaluated in chunks)."""
    out = np.empty(len(values), dtype=np.int64)
    for a in range(0, len(values), chunk):
        v = values[a:a + chunk]
        out[a:a + len(v)] = np.abs(table[None, :] - v[:, None]).argmin(axis=1)
    return out


# ---- tables -----------------------------------------------------------------------------------------------
def pec_block(element, wavelength, ion_digit, transition, root=DEFAULT_INPUT):
    directory = os.path.join(root, "PEC", element)
    tag = element.lower() + str(ion_digit)
    path = [os.path.join(directory, f) for f in sorted(os.listdir(directory)) if tag in f][0]
    with open(path) as fh:
        lines = fh.readlines()
    head = None
    for idx, line in enumerate(lines):
        items = line.split()
        try:
            if float(items[0]) == wavelength and transition in items:
                head, head_idx = items, idx + 1
                break
        except (ValueError, IndexError):
            break
    n_ne, n_te = int(head[2]), int(head[3])
    n_rows = int(np.ceil(n_ne / 8) + np.ceil(n_te / 8) + n_ne * np.ceil(n_te / 8))
    vals = np.array([v for line in islice(lines, head_idx, head_idx + n_rows) for v in line.split()], dtype=np.float64)
    return vals[:n_ne], vals[n_ne:n_ne + n_te], vals[n_ne + n_te:].reshape(n_ne, n_te)


def pec_table(element, wavelength, ion_digit, transition, interp_step=2000, root=DEFAULT_INPUT):
    """(ne_grid, te_grid, pec[interp_step, interp_step]) as pec.py:153-187 builds them."""
    from scipy.interpolate import RectBivariateSpline

    ne_nodes, te_nodes, table = pec_block(element, wavelength, ion_digit, transition, root)
    ne_new = np.linspace(ne_nodes[0], ne_nodes[-1], interp_step, endpoint=True)
    te_new = np.linspace(te_nodes[0], te_nodes[-1], interp_step, endpoint=True)
    spl = RectBivariateSpline(ne_nodes, te_nodes, table, kx=1, ky=1, s=0)
    return ne_new, te_new, spl(ne_new, te_new)


def fa_table(element, root=DEFAULT_INPUT):
    """(T grid, {column: values}) as fractional_abundance.py:59-91."""
    from scipy.interpolate import interp1d

    import pandas as pd

    # pandas' CSV float parser (not correctly rounded in the last ulp) is part of the reference's arithmetic
    path = os.path.join(root, "fractional_abundance", f"fractional_abundance_{element}.dat")
    df = pd.read_csv(path, delimiter=" ")
    names = list(df.columns)
    temp = df["T"]
    Te = np.arange(temp.iloc[0], temp.iloc[-1], 1)
    cols = {name: interp1d(temp.iloc[:], df[name])(Te) / 100 for name in names[1:]}
    return Te, cols, names


def two_gauss_profile(ne_coeff, te_coeff, max_reff=0.539):
    reff = np.arange(0, max_reff, 0.001)

    def eq(c):
        A1, x1, w1, A2, x2, w2 = c
        return A1 * np.exp(-((reff - x1) ** 2) / (2 * w1**2)) + A2 * np.exp(-((reff - x2) ** 2) / (2 * w2**2))

    return np.column_stack((reff, eq(te_coeff), eq(ne_coeff)))  # Reff, T_e, n_e


def read_observed_volume(element, root=DEFAULT_INPUT):
    import pandas as pd

    path = os.path.join(root, "geometry", "observed_plasma_volume", element,
                        f"{element}_plasma_coordinates-10_mm_spacing-height_30-length_20-slit_100.dat")
    return pd.read_csv(path, sep=";").to_numpy(dtype=float)  # reader.py:150-158


def read_reff(name, root=DEFAULT_INPUT):
    import pandas as pd

    path = os.path.join(root, "geometry", "reff", f"{name}.dat")
    return pd.read_csv(path, sep=";").iloc[:, 4].to_numpy(dtype=float)  # reader.py:116-128


# ---- the pipeline -------------------------------------------------------------------------------------------
def emissivity(element, ion_state, wavelength, impurity_concentration, transitions, profile, observed, reff_all,
               root=DEFAULT_INPUT, tables=None):
    """*profile* is (K,3) [Reff, T_e, n_e]; *observed* the (V,5) Stage-1 table; *reff_all* Reff per lattice voxel.

    Returns a namespace with ``rows`` (V'', 15) in the reference's column order (rows in Stage-1 file order, i.e.
    before the reference's final sort by Reff), ``flux`` {transition: value, "TOTAL": value}, ``reff_boundary``.
    """
    conc = impurity_concentration / 100  # reader.py:132-134
    idx = observed[:, 0].astype(np.int64)
    reff = reff_all[idx]
    keep = ~np.isnan(reff)  # dropna, reader.py:176
    obs, reff = observed[keep], reff[keep]

    boundary = min(reff.max(), profile[:, 0].max())  # reader.py:210-217
    k = nearest_rows(profile[:, 0], reff)
    ne, te = profile[k, 2], profile[k, 1]
    inside = ~(reff >= boundary)  # reader.py:240-242
    obs, reff, ne, te = obs[inside], reff[inside], ne[inside], te[inside]

    if tables is None:
        tables = load_tables(element, ion_state, wavelength, transitions, root)
    j = nearest_rows(tables.fa_T, te)
    fa_ion, fa_stripped = tables.fa_ion[j], tables.fa_last[j]

    pec = {}
    for tr in transitions:
        ne_grid, te_grid, tab = tables.pec[tr]
        pec[tr] = tab[nearest_rows(ne_grid, ne), nearest_rows(te_grid, te)]
    emis = {}
    for tr in transitions:
        if "RECOM" in tr:
            emis[tr] = fa_stripped * pec[tr] * ne**2 * conc
        else:
            emis[tr] = fa_ion * pec[tr] * ne**2 * conc
    total = np.sum(np.column_stack([emis[tr] for tr in transitions]), axis=1)
    w = obs[:, 4]
    flux = {tr: float((emis[tr] * w).sum()) for tr in transitions}
    flux["TOTAL"] = float((total * w).sum())
    cols = [obs[:, 0], obs[:, 1], obs[:, 2], obs[:, 3], w, reff, ne, te, fa_ion, fa_stripped]
    cols += [pec[tr] for tr in transitions] + [emis[tr] for tr in transitions] + [total]
    return SimpleNamespace(rows=np.column_stack(cols), flux=flux, reff_boundary=float(boundary),
                           total_str=f"{flux['TOTAL']:.2e}")


def load_tables(element, ion_state, wavelength, transitions, root=DEFAULT_INPUT):
    T, cols, names = fa_table(element, root)
    return SimpleNamespace(
        fa_T=T, fa_ion=cols[ion_state], fa_last=cols[names[-1]],
        pec={tr: pec_table(element, wavelength, ion_state[-1], tr, 2000, root) for tr in transitions},
    )
