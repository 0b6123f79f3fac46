"""The weight-only kernels find m = rint(100 * degrees(arccos(c))) - the reference's round(angle, 2) of simulation.py:520-531 -
without the fp64 arc cosine: m is the index with cos_bound[m + 1] < c <= cos_bound[m], cos_bound[j] = cos((j - 0.5) / 100 deg)
(visibility.cu: certain_ray_weight; table built in geometry_host.cu).  This restates the rule in NumPy and checks it against
the reference's expression on cosines spread over the whole range and clustered around the rounding boundaries: the two agree
everywhere except within 1e-9 (in units of 0.01 degree) of a boundary - the rays the kernels count as near_rounding_rays."""
import numpy as np


def m_reference(c):
    return np.rint(np.degrees(np.arccos(c)) * 100.0)  # np.round(x, 2) is rint(x * 100) / 100


def m_table(c, cos_bound):
    # cos_bound decreases with j: count the boundaries that are >= c, minus one
    asc = cos_bound[::-1]
    n_ge = len(asc) - np.searchsorted(asc, c, side="left")
    return (n_ge - 1).astype(float)


def test_table_rule_equals_rounding_the_angle():
    j = np.arange(18002, dtype=np.float64)
    cos_bound = np.cos((j - 0.5) * (np.pi / 18000.0))
    cos_bound[0], cos_bound[-1] = 2.0, -2.0  # the angle lies in [0, 180] degrees: no boundary beyond
    assert np.all(np.diff(cos_bound) < 0)
    rng = np.random.default_rng(0)
    broad = np.cos(np.radians(rng.uniform(0.01, 179.99, 2_000_000)))
    # clustered around boundaries: angles (m + 0.5 + delta) / 100 with |delta| from 1e-12 to 1e-3
    m0 = rng.integers(100, 17900, 400_000).astype(np.float64)
    delta = rng.choice([-1.0, 1.0], 400_000) * 10.0 ** rng.uniform(-12, -3, 400_000)
    near = np.cos(np.radians((m0 + 0.5 + delta) / 100.0))
    for c in (broad, near):
        ref, tab = m_reference(c), m_table(c, cos_bound)
        deg100 = np.degrees(np.arccos(c)) * 100.0
        dist = np.abs(np.abs(deg100 - np.rint(deg100)) - 0.5)  # distance of the angle to its nearest rounding boundary
        differ = ref != tab
        assert not np.any(differ & (dist >= 1e-9)), (c[differ & (dist >= 1e-9)][:5], dist[differ][:5])
        assert np.all(np.abs(ref[differ] - tab[differ]) == 1)
    ends = np.array([1.0, -1.0, np.cos(np.radians(0.004)), np.cos(np.radians(179.996))])
    assert np.array_equal(m_reference(ends), m_table(ends, cos_bound))
    # the far-from-boundary majority is decided identically
    assert np.mean(m_reference(broad) == m_table(broad, cos_bound)) > 0.999999
