"""Synthetic grids / matrices have the shapes BASELINE.json names."""

import numpy as np
import pytest
from scipy.sparse import csr_array

from anemoi_transform_b200 import synthetic as syn
from anemoi_transform_b200.grids import lookup


def test_o96_matches_reference_pins():
    # reference tests/test_grids.py:49-57
    x = lookup("o96")
    assert x["latitudes"].shape == (40320,) and x["longitudes"].shape == (40320,)
    assert x["latitudes"].mean() == pytest.approx(0.0)
    assert x["longitudes"].mean() == pytest.approx(179.14285714285714)
    assert x["latitudes"][31415] == pytest.approx(-31.324557701757268)
    assert x["longitudes"][31415] == pytest.approx(224.32835820895522)


def test_grid_sizes():
    assert syn.regular_latlon(1.0)[0].size == 65_160
    assert syn.regular_latlon(0.25)[0].size == 1_038_240
    lat, lon = syn.n320_like()
    assert lat.size == lon.size == 542_080 and np.unique(lat).size == 640
    assert lon.min() >= 0 and lon.max() < 360
    la, lo = syn.rotated_lam(10, 12, 0.02)
    assert la.size == 120 and abs(la.mean() - 60.0) < 0.1 and abs(lo.mean() - 10.0) < 0.2


def test_bilinear_matrix_interpolates_linear_fields_and_writes_the_npz_schema(tmp_path):
    s_lat, s_lon = syn.regular_latlon(1.0)
    t_lat, t_lon = syn.octahedral(96)
    d, i, p, shape = syn.bilinear_matrix(1.0, t_lat, t_lon)
    assert shape == (40_320, 65_160) and d.dtype == np.float32 and i.dtype == np.int32 and d.size == 161_280
    m = csr_array((d, i, p), shape=shape)
    assert np.allclose(m @ s_lat, t_lat, atol=1e-4)
    assert np.allclose(np.asarray(m.sum(axis=1)).ravel(), 1.0, atol=1e-6)
    assert (np.diff(i.reshape(-1, 4), axis=1) > 0).all()  # columns sorted within each row
    syn.save_regrid_npz(tmp_path / "m.npz", d, i, p, shape, s_lat, s_lon, t_lat, t_lon)
    z = np.load(tmp_path / "m.npz")
    # make-regrid-file.py:150-160
    assert set(z.keys()) == {"matrix_data", "matrix_indices", "matrix_indptr", "matrix_shape", "in_latitudes", "in_longitudes", "out_latitudes", "out_longitudes"}
