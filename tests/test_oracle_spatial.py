"""The spatial oracle against (a) the golden outputs of the imported reference, (b) the
reference's own known-answer tests (tests/test_spatial.py) and (c) an exhaustive search that
spells out cKDTree's distance arithmetic."""

import numpy as np
import pytest
from scipy.spatial import cKDTree

from oracle import spatial as osp


def _lam11():
    la, lo = np.meshgrid(np.linspace(44.0, 46.0, 11), np.linspace(0.0, 2.0, 11))
    return la.flatten(), lo.flatten()


# ---- the reference's known-answer tests, tests/test_spatial.py:18-145 ----------------------
@pytest.mark.parametrize("cropping_distance", [1.0, 3.0, 5.0])
def test_kat_cutout_with_max_distance(cropping_distance):
    lam_lats, lam_lons = _lam11()
    g_lats = np.array([43.1, 44.0, 45.0, 45.5, 46.0, 50.0])
    g_lons = np.array([359.1, 359.5, 0.0, 1.0, 2.0, 0.0])
    for fn in (osp.cutout_mask, osp.cutout_mask_vectorised):
        mask = fn(lam_lats, lam_lons, g_lats, g_lons, cropping_distance=cropping_distance, max_distance_km=250.0)
        assert np.array_equal(mask, [True, False, False, False, False, False])


def test_kat_cutout_with_min_distance():
    lam_lats, lam_lons = _lam11()
    g_lats = np.array([44.0, 45.0, 46.0, 46.1, 47.5])
    g_lons = np.array([0.0, 1.0, 2.0, -0.1, -1.5])
    for fn in (osp.cutout_mask, osp.cutout_mask_vectorised):
        assert np.array_equal(fn(lam_lats, lam_lons, g_lats, g_lons, min_distance_km=100.0), [False, False, False, False, True])


def test_kat_cutout_large_grid():
    la, lo = np.meshgrid(np.linspace(40.0, 50.0, 21), np.linspace(0.0, 10.0, 21))
    gla, glo = np.meshgrid(np.linspace(30.0, 60.0, 31), np.linspace(-10.0, 20.0, 31))
    mask = osp.cutout_mask_vectorised(la.flatten(), lo.flatten(), gla.flatten(), glo.flatten(), min_distance_km=150.0, max_distance_km=300.0)
    assert mask.shape == (961,) and mask.dtype == bool and mask.any() and not mask.all()
    assert np.array_equal(mask, osp.cutout_mask(la.flatten(), lo.flatten(), gla.flatten(), glo.flatten(), min_distance_km=150.0, max_distance_km=300.0))


# ---- golden outputs of the imported reference ------------------------------------------------
def test_golden_latlon_xyz(golden_spatial):
    g = golden_spatial
    x, y, z = osp.latlon_to_xyz(g["o_lat"], g["o_lon"])
    assert np.array_equal(x, g["o_x"]) and np.array_equal(y, g["o_y"]) and np.array_equal(z, g["o_z"])
    la, lo = osp.xyz_to_latlon(x, y, z)
    assert np.array_equal(la, g["o_lat_back"]) and np.array_equal(lo, g["o_lon_back"])


def test_golden_cropping(golden_spatial):
    g = golden_spatial
    assert np.array_equal(osp.cropping_mask(g["g_lat"], g["g_lon"], 70.0, -20.0, 40.0, 15.0), g["crop_wrap"])
    assert np.array_equal(osp.cropping_mask(g["g_lat"], g["g_lon"] - 360.0, 10.0, 100.0, -10.0, 140.0), g["crop_plus360"])
    assert g["crop_wrap"].sum() > 0 and g["crop_plus360"].sum() > 0


def test_golden_knn(golden_spatial):
    g = golden_spatial
    assert np.array_equal(osp.nearest_grid_points(g["g_lat"], g["g_lon"], g["o_lat"], g["o_lon"]), g["ngp_k1"])
    i4, d4 = osp.nearest_grid_points(g["g_lat"], g["g_lon"], g["lam_lat"], g["lam_lon"], num_neighbours_to_return=4, return_distances=True)
    assert np.array_equal(i4, g["ngp_k4_idx"]) and np.array_equal(d4, g["ngp_k4_dist"])
    iu, du = osp.nearest_grid_points(g["lam_lat"], g["lam_lon"], g["o_lat"], g["o_lon"], max_distance=0.01, return_distances=True)
    assert np.array_equal(iu, g["ngp_ub_idx"]) and np.array_equal(du, g["ngp_ub_dist"])
    assert (iu == g["lam_lat"].size).any() and (iu < g["lam_lat"].size).any()  # misses and hits


def test_golden_masks(golden_spatial):
    g = golden_spatial
    lam, o = (g["lam_lat"], g["lam_lon"]), (g["o_lat"], g["o_lon"])
    assert np.array_equal(osp.thinning_mask(*lam, *o), g["thinning"])
    assert np.array_equal(osp.thinning_mask(*lam, g["g_lat"], g["g_lon"], cropping_distance=6.0), g["thinning_crop6"])
    assert np.array_equal(osp.global_on_lam_mask(*lam, *o), g["gol_none"])
    assert np.array_equal(osp.global_on_lam_mask(*lam, *o, distance_km=150.0), g["gol_150km"])
    empty = osp.global_on_lam_mask(*lam, *o, distance_km=1.0)
    assert empty.shape == (0,) and empty.dtype == g["gol_1km_empty"].dtype == np.float64


@pytest.mark.parametrize("fn", [osp.cutout_mask, osp.cutout_mask_vectorised])
def test_golden_cutout(golden_spatial, fn):
    g = golden_spatial
    lam, o = (g["lam_lat"], g["lam_lon"]), (g["o_lat"], g["o_lon"])
    assert np.array_equal(fn(*lam, *o), g["cutout_default"])
    assert np.array_equal(fn(*lam, *o, min_distance_km=80.0, max_distance_km=400.0), g["cutout_min80_max400"])
    assert np.array_equal(fn(*lam, *o, cropping_distance=5.0, neighbours=3, min_distance_km=10), g["cutout_n3_crop5"])
    assert np.array_equal(fn(*lam, g["g_lat"], g["g_lon"]), g["cutout_regular_default"])
    assert 0 < (~g["cutout_default"]).sum() < g["cutout_default"].size


# ---- what cKDTree computes --------------------------------------------------------------------
def test_ckdtree_distance_is_unfused_float64_sum_of_squares(golden_spatial):
    g = golden_spatial
    src = np.array(osp.latlon_to_xyz(g["g_lat"], g["g_lon"])).T
    tgt = np.array(osp.latlon_to_xyz(g["lam_lat"], g["lam_lon"])).T
    d, i = cKDTree(src).query(tgt, k=3)
    bi, bd, tie = osp.knn_bruteforce(src, tgt, k=3)
    assert np.array_equal(d, bd)  # bitwise
    differ = (i != bi).any(axis=1)
    assert not (differ & ~tie).any()  # indices differ only where d² ties exist
    # strict upper bound: d < bound
    dmin = d[:, 0].min()
    du, iu = cKDTree(src).query(tgt, k=1, distance_upper_bound=dmin)
    assert (iu == src.shape[0]).all() and np.isinf(du).all()
    # ball query is d² <= r·r
    r = float(np.median(d[:, 1]))
    balls = cKDTree(src).query_ball_point(tgt[:50], r)
    for t, b in zip(tgt[:50], balls):
        diff = src - t
        d2 = (diff[:, 0] * diff[:, 0] + diff[:, 1] * diff[:, 1]) + diff[:, 2] * diff[:, 2]
        assert set(b) == set(np.nonzero(d2 <= r * r)[0])


def test_ckdtree_pads_when_k_exceeds_sources():
    src = np.array([[1.0, 0, 0], [0, 1.0, 0]])
    d, i = cKDTree(src).query(np.array([[1.0, 0, 0]]), k=3)
    assert i[0, 2] == 2 and np.isinf(d[0, 2])


def test_outline_restatement_matches_the_reference_golden(golden_spatial):
    """oracle.spatial.outline (spatial.py:539-584) against the imported reference's outputs."""
    g = golden_spatial
    assert osp.outline(g["patch_lat"], g["patch_lon"]) == g["outline_patch"].tolist()
    assert osp.outline(g["patch_lat"], g["patch_lon"], neighbours=7) == g["outline_patch_n7"].tolist()
    assert osp.outline(g["lam_lat"], g["lam_lon"]) == g["outline_lam"].tolist()
