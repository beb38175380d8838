"""Oracle: numpy restatement of the earthkit-meteo functions behind the pointwise filters.

TEST INFRASTRUCTURE — see oracle/__init__.py.

The reference delegates this arithmetic to earthkit-meteo (`pyproject.toml:40`,
`earthkit-meteo>=0.4.1,<1`, no lock file), which is neither vendored under /root/reference
nor installed here.  The functions below restate its published algorithms:

    xy_to_polar / polar_to_xy                       earthkit.meteo.wind.array
        called at  filters/fields/uv_to_ddff.py:94-98, 120-124
    relative_humidity_from_specific_humidity,
    specific_humidity_from_relative_humidity,
    saturation_vapour_pressure (mixed phase),
    vapour_pressure_from_specific_humidity,
    specific_humidity_from_vapour_pressure           earthkit.meteo.thermo.array
        called at  filters/fields/q_to_r.py:72, 78-80

PINNED by the reference's own golden vectors for this boundary (np.allclose, rtol 1e-5):
tests/field_filters/test_uv_to_ddff.py:24-42 and
tests/field_filters/test_pressure_level_humidity.py:27-40 — checked in
tests/test_oracle_pointwise.py.  Bitwise agreement with earthkit-meteo itself is unpinned
(the package is unavailable); the stated tolerance of the path is 1e-6 of the field range.

numpy keeps float32 inputs float32 (Python-float constants are weak scalars, NEP 50), so
these functions compute in the dtype of their inputs, like the library does.
"""

from __future__ import annotations

import numpy as np

# earthkit.meteo.constants.constants
R_earth = 6371229.0
radian = np.pi / 180.0
degree = 180.0 / np.pi
Rd = 287.0597
Rv = 461.5250
epsilon = Rd / Rv
T0 = 273.16


def direction(u, v, convention="meteo", to_positive=True):
    """Wind direction in degrees; "meteo": direction the wind blows FROM, clockwise from north."""
    if convention != "meteo":
        raise NotImplementedError(convention)
    minus_pi2 = -np.pi / 2.0
    d = np.arctan2(v, u)
    d = np.asarray(d)
    m = d <= minus_pi2
    out = np.empty_like(d)
    out[m] = (minus_pi2 - d[m]) * degree
    m = ~m
    out[m] = (1.5 * np.pi - d[m]) * degree
    return out


def xy_to_polar(x, y, convention="meteo"):
    """(u, v) → (speed, direction)."""
    return np.hypot(x, y), direction(x, y, convention=convention)


def polar_to_xy(magnitude, direction, convention="meteo"):
    """(speed, direction) → (u, v)."""
    if convention != "meteo":
        raise NotImplementedError(convention)
    a = (270.0 - direction) * radian
    return magnitude * np.cos(a), magnitude * np.sin(a)


def _es_water(t):
    return 611.21 * np.exp(17.502 * (t - T0) / (t - 32.19))


def _es_ice(t):
    return 611.21 * np.exp(22.587 * (t - T0) / (t + 0.7))


def saturation_vapour_pressure(t):
    """Mixed-phase saturation vapour pressure (IFS): ice ≤ 250.16 K, water ≥ 273.16 K."""
    t = np.asarray(t)
    ti = T0 - 23.0
    svp = np.empty_like(t)
    i_mask = t <= ti
    w_mask = t >= T0
    m_mask = ~(i_mask | w_mask)
    with np.errstate(over="ignore", invalid="ignore", divide="ignore"):
        svp[i_mask] = _es_ice(t[i_mask])
        svp[w_mask] = _es_water(t[w_mask])
        tm = t[m_mask]
        alpha = np.square(tm - ti) / np.square(T0 - ti)
        svp[m_mask] = alpha * _es_water(tm) + (1.0 - alpha) * _es_ice(tm)
    return svp


def vapour_pressure_from_specific_humidity(q, p):
    c = epsilon * (1.0 / epsilon - 1.0)
    return (p * q) / (epsilon + c * q)


def specific_humidity_from_vapour_pressure(e, p, eps=1e-4):
    v = np.asarray(p + (epsilon - 1.0) * e)
    v = np.array(v, copy=True)
    v[np.asarray(p - e) < eps] = np.nan
    return epsilon * e / v


def relative_humidity_from_specific_humidity(t, q, p):
    svp = saturation_vapour_pressure(t)
    e = vapour_pressure_from_specific_humidity(q, p)
    with np.errstate(over="ignore", invalid="ignore", divide="ignore"):
        return 100.0 * e / svp


def specific_humidity_from_relative_humidity(t, r, p):
    svp = saturation_vapour_pressure(t)
    with np.errstate(over="ignore", invalid="ignore", divide="ignore"):
        e = r * svp / 100.0
        return specific_humidity_from_vapour_pressure(e, p)


def saturation_vapour_pressure_water(t):
    """Water-phase Tetens formula (earthkit.meteo.thermo.array.saturation_vapour_pressure(t, phase="water"))."""
    return _es_water(np.asarray(t))


def temperature_from_saturation_vapour_pressure(es):
    """Inverse of the water-phase formula (earthkit.meteo.thermo.array)."""
    v = np.log(es / 611.21)
    return (v * 32.19 - 17.502 * T0) / (v - 17.502)


def dewpoint_from_relative_humidity(t, r):
    """earthkit.meteo.thermo.dewpoint_from_relative_humidity, called at dewpoint.py:65.
    PINNED by tests/field_filters/test_dewpoint.py:23-27 (tests/test_oracle_pointwise.py)."""
    with np.errstate(over="ignore", invalid="ignore", divide="ignore"):
        es = saturation_vapour_pressure_water(t) * r / 100.0
        return temperature_from_saturation_vapour_pressure(es)


def relative_humidity_from_dewpoint(t, td):
    """earthkit.meteo.thermo.relative_humidity_from_dewpoint, called at dewpoint.py:73."""
    with np.errstate(over="ignore", invalid="ignore", divide="ignore"):
        return 100.0 * saturation_vapour_pressure_water(td) / saturation_vapour_pressure_water(t)


def clip(data, minimum, maximum):
    """clipper.py:69."""
    return np.clip(data, minimum, maximum)


def apply_mask(values, mask):
    """apply_mask.py:184-185 on a flattened copy."""
    values = np.array(values, copy=True).reshape(-1)
    values[mask] = np.nan
    return values
