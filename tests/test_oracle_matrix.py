"""The numpy restatement of the bilinear matrix builder (oracle/matrix.py), pinned by the
properties that define the scheme — the reference has no golden vector for matrix
construction (it is MIR's job there)."""

import numpy as np

from anemoi_transform_b200 import synthetic as syn
from oracle import matrix as om


def test_regular_grid_detection():
    lat, lon = syn.regular_latlon(2.0)
    assert om.regular_grid_parameters(lat, lon) == (90.0, -2.0, 91, 0.0, 2.0, 180)
    assert om.regular_grid_parameters(*syn.octahedral(16)) is None
    assert om.regular_grid_parameters(lat[:-1], lon[:-1]) is None
    # a regional (non-periodic) grid is not accepted: the builder wraps in longitude
    sub = (lon < 100.0)
    assert om.regular_grid_parameters(lat[sub], lon[sub]) is None


def test_bilinear_rows_sum_to_one_and_reproduce_linear_fields():
    s_lat, s_lon = syn.regular_latlon(1.0)
    t_lat, t_lon = syn.octahedral(32)
    m = om.bilinear_csr(s_lat, s_lon, t_lat, t_lon)
    assert m.shape == (t_lat.size, s_lat.size) and m.nnz == 4 * t_lat.size
    assert np.allclose(np.asarray(m.sum(axis=1)).ravel(), 1.0, atol=2e-7)
    assert (m.data >= 0).all() and (np.diff(m.indices.reshape(-1, 4), axis=1) > 0).all()
    # linear in latitude everywhere; linear in longitude away from the 360 -> 0 seam
    assert np.allclose(m @ s_lat, t_lat, atol=1e-4)
    inner = t_lon < 359.0
    assert np.allclose((m @ s_lon)[inner], t_lon[inner], atol=1e-3)
    # a smooth periodic field is interpolated to second order
    f = lambda la, lo: np.cos(np.deg2rad(la)) * np.sin(2 * np.deg2rad(lo))  # noqa: E731
    assert np.abs(m @ f(s_lat, s_lon) - f(t_lat, t_lon)).max() < 3e-4


def test_target_on_a_source_point_takes_its_value():
    s_lat, s_lon = syn.regular_latlon(2.0)
    pick = np.array([0, 181, 5000, s_lat.size - 1])
    m = om.bilinear_csr(s_lat, s_lon, s_lat[pick], s_lon[pick])
    v = np.random.default_rng(0).standard_normal(s_lat.size)
    assert np.array_equal(m @ v, v[pick])


def test_matches_the_bench_matrix():
    """The same scheme as synthetic.bilinear_matrix, which every config of the bench uses."""
    t_lat, t_lon = syn.octahedral(24)
    d, i, p, shape = syn.bilinear_matrix(2.0, t_lat, t_lon)
    d2, i2, p2, shape2 = om.bilinear_matrix(90.0, -2.0, 91, 0.0, 2.0, 180, t_lat, t_lon)
    assert shape == shape2 and np.array_equal(i, i2) and np.array_equal(p, p2) and np.array_equal(d, d2)
