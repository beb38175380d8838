"""Oracle restatements against the UNMODIFIED reference imported from /root/reference —
build container only (the GPU box has no reference tree; the same comparisons are frozen in
tests/golden/ by oracle/make_golden.py)."""

import numpy as np
import pytest

from oracle import reference_import
from oracle import spatial as osp

pytestmark = [pytest.mark.reference, pytest.mark.skipif(not reference_import.available(), reason="reference sources not present")]


@pytest.fixture(scope="module")
def ref():
    return reference_import.load()


def test_spatial_functions_agree_on_random_inputs(ref):
    from anemoi_transform_b200 import synthetic as syn

    sp = ref["spatial"]
    for seed in range(3):
        rng = np.random.default_rng(seed)
        lam = syn.rotated_lam(12 + seed, 15, 0.7, float(rng.uniform(-60, 70)), float(rng.uniform(0, 360)))
        glob = syn.octahedral(12 + 4 * seed)
        assert np.array_equal(sp.cutout_mask(*lam, *glob), osp.cutout_mask_vectorised(*lam, *glob))
        assert np.array_equal(sp.cutout_mask(*lam, *glob, min_distance_km=50, max_distance_km=900), osp.cutout_mask(*lam, *glob, min_distance_km=50, max_distance_km=900))
        assert np.array_equal(sp.thinning_mask(*lam, *glob), osp.thinning_mask(*lam, *glob))
        assert np.array_equal(sp.global_on_lam_mask(*lam, *glob), osp.global_on_lam_mask(*lam, *glob))
        assert np.array_equal(sp.nearest_grid_points(*glob, *lam, num_neighbours_to_return=3), osp.nearest_grid_points(*glob, *lam, num_neighbours_to_return=3))


def test_goldens_are_reproducible(ref, golden_spatial):
    g = golden_spatial
    sp = ref["spatial"]
    assert np.array_equal(sp.cutout_mask(g["lam_lat"], g["lam_lon"], g["o_lat"], g["o_lon"]), g["cutout_default"])
    assert np.array_equal(sp.nearest_grid_points(g["g_lat"], g["g_lon"], g["o_lat"], g["o_lon"]), g["ngp_k1"])


def test_patch_data_request_matches_the_reference_filters(ref):
    """Host-only behaviour of the product's filters against the reference's classes: the
    `patch_data_request` rewrites (no GPU involved)."""
    import copy

    from anemoi_transform_b200.filters import create_filter_by_name as F

    cases = [
        (F("orog_to_z_fields"), ref["orog_to_z"].Orography() if "orog_to_z" in ref else None,
         [{"param": ["z", "t"], "levtype": "pl"}, {"param": ["orog", "t"], "levelist": [500]}, {"param": ["z", "t"], "levtype": "sfc"}, {"param": "z", "levtype": "pl"}, {"param": ["q"]}, {}]),
        (F("lnsp_to_sp"), ref["lnsp_to_sp"].LnspToSp(), [{"param": ["sp", "t"]}, {"param": ["lnsp"]}, {"param": ["q"]}, {}]),
        (F("cos_sin_from_rad", param="x"), ref["cos_sin_from_rad"].CosSinFromRad(param="x"), [{"param": ["cos_x", "t"]}, {"param": ["sin_x", "cos_x"]}, {"param": ["q"]}, {}]),
        (F("cos_sin_mean_wave_direction"), ref["cos_sin_mean_wave_direction"].CosSinWaveDirection(), [{"param": ["cos_mwd", "t"]}, {"param": ["mwd"]}, {}]),
    ]  # fmt: skip
    for mine, theirs, requests in cases:
        if theirs is None:
            continue
        for req in requests:
            assert mine.patch_data_request(copy.deepcopy(req)) == theirs.patch_data_request(copy.deepcopy(req)), (type(mine).__name__, req)
    for both, flt in (({"param": ["sp", "lnsp"]}, F("lnsp_to_sp")), ({"param": ["z", "orog"]}, F("orog_to_z_fields"))):
        with pytest.raises(ValueError, match="cannot contain both"):
            flt.patch_data_request(both)
