"""SURVEY §8(f) rank 2 and the round-2 planner extensions:

* `at_bilinear_matrix` (device matrix construction) against the numpy restatement, bit for bit;
* `regrid(in_grid=…, out_grid=…, method=…)` recipes running offline on locally built matrices;
* pipelines whose regrid is a nearest-neighbour / mask gather, and dewpoint / cos-sin followers,
  fused into one launch and bitwise equal to the filters run one after the other.
"""

import numpy as np
import pytest
from conftest import assert_same_values
from scipy.sparse import csr_array

from anemoi_transform_b200 import ekd
from anemoi_transform_b200 import synthetic as syn
from oracle import matrix as om

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def F(cuda):
    from anemoi_transform_b200.filters import create_filter_by_name

    return create_filter_by_name


@pytest.mark.parametrize("step,target", [(2.0, "o48"), (1.0, "o96"), (0.25, "n320")])
def test_device_bilinear_matrix_is_bitwise_the_scipy_built_one(cuda, step, target):
    from anemoi_transform_b200.regrid_files import make_bilinear_matrix

    s_lat, s_lon = syn.regular_latlon(step)
    t_lat, t_lon = syn.n320_like() if target == "n320" else syn.octahedral(int(target[1:]))
    # targets on source lines and on the 360 -> 0 seam exercise the explicit zeros and the wrap
    t_lat, t_lon = np.concatenate([t_lat, [90.0, -90.0, 0.0, 10.0]]), np.concatenate([t_lon, [0.0, 359.9999, 360.0 - step / 2, 720.5]])
    d, i, p, shape = make_bilinear_matrix(s_lat, s_lon, t_lat, t_lon)
    want = om.bilinear_matrix(*om.regular_grid_parameters(s_lat, s_lon), t_lat, t_lon)
    assert shape == want[3] and d.dtype == np.float32 and i.dtype == np.int32
    assert np.array_equal(i, want[1]) and np.array_equal(p, want[2])
    assert np.array_equal(d.view(np.uint32), want[0].view(np.uint32)), "weights are not bitwise equal"
    assert (d == 0).any()  # the explicit zeros are kept


def test_bilinear_builder_refuses_what_it_cannot_build(cuda):
    from anemoi_transform_b200.regrid_files import make_bilinear_matrix

    t_lat, t_lon = syn.octahedral(16)
    with pytest.raises(NotImplementedError, match="regular"):
        make_bilinear_matrix(*syn.octahedral(32), t_lat, t_lon)
    s_lat, s_lon = syn.regular_latlon(2.0)
    keep = s_lat <= 60.0
    with pytest.raises(ValueError, match="leave the source grid"):
        make_bilinear_matrix(s_lat[keep], s_lon[keep], t_lat, t_lon)


def _fieldlist(values, lat, lon, specs):
    return ekd.from_source("list-of-dicts", [dict(values=v, latitudes=lat, longitudes=lon, **s) for v, s in zip(values, specs)])


def test_in_grid_out_grid_method_recipes_run_offline(F, tmp_path):
    """regrid.py:211-259, 455-467: `in_grid` / `out_grid` / `method` — linear on a locally built
    bilinear matrix, nearest-neighbour on a k = 1 matrix; results equal the scipy matrix of the
    same construction applied per field."""
    s_lat, s_lon = syn.regular_latlon(2.0)
    t_lat, t_lon = syn.octahedral(32)
    values = [syn.synthetic_field("t", s_lat.size, s, 0.002 if s == 2 else 0.0) for s in range(9)]
    data = _fieldlist(values, s_lat, s_lon, [dict(param="t", levelist=850, step=k) for k in range(9)])
    m = om.bilinear_csr(s_lat, s_lon, t_lat, t_lon)
    for kwargs in (dict(in_grid=[2.0, 2.0], out_grid="O32"), dict(out_grid="O32"), dict(in_grid=[2.0, 2.0], out_grid="O32", method="linear"), dict(out_grid=dict(latitudes=t_lat, longitudes=t_lon))):
        out = F("regrid", **kwargs).forward(data)
        assert len(out) == 9
        for k, f in enumerate(out):
            assert_same_values(f.to_numpy(), m @ values[k], f"linear {kwargs} field {k}")
            assert np.array_equal(f.grid_points()[0], t_lat) and np.array_equal(f.grid_points()[1], t_lon)
    from scipy.spatial import cKDTree

    from oracle import spatial as osp

    out = F("regrid", in_grid=[2.0, 2.0], out_grid="O32", method="nearest-neighbour").forward(data)
    dist, idx = cKDTree(np.array(osp.latlon_to_xyz(s_lat, s_lon)).T).query(np.array(osp.latlon_to_xyz(t_lat, t_lon)).T, k=1)
    for k, f in enumerate(out):
        got, want = f.to_numpy(), values[k][idx]
        agree = (got == want) | (np.isnan(got) & np.isnan(want))
        assert agree.mean() > 0.995  # exact-distance ties may take the other source
    with pytest.raises(NotImplementedError, match="not built locally"):
        F("regrid", in_grid=[2.0, 2.0], out_grid="O32", method="grid-box-average")
    with pytest.raises(NotImplementedError, match="regular"):
        F("regrid", in_grid="O48", out_grid="O32").forward(_fieldlist([np.zeros(syn.octahedral(48)[0].size, np.float32)], *syn.octahedral(48), [dict(param="t", levelist=1)]))


def _mixed_fields(s_lat, s_lon):
    specs = [("t", 850), ("u", 850), ("z", 500), ("v", 850), ("rh", 850), ("u", 500), ("q", 850), ("v", 500), ("t", 500), ("rh", 500), ("q", 500), ("lsm", 0), ("cos_mwd", 0), ("sin_mwd", 0), ("t2", 850), ("t2", 500)]
    values = []
    for i, (p, _) in enumerate(specs):
        if p == "t2":
            values.append(syn.synthetic_field("t", s_lat.size, 60 + i))
            continue
        if p == "rh":
            v = np.random.default_rng(i).uniform(0.0, 100.0, s_lat.size).astype(np.float32)
            v[::97] = 0.0  # r == 0 is replaced by 1e-4 inside the dewpoint conversion
        elif p in ("cos_mwd", "sin_mwd"):
            v = np.random.default_rng(i).uniform(-1.0, 1.0, s_lat.size).astype(np.float32)
        else:
            v = syn.synthetic_field(p, s_lat.size, 60 + i, 0.002 if i % 4 == 0 else 0.0)
        if p == "u":
            v[5::211] = -0.0  # a gather must keep the sign of zero
        values.append(v)
    return _fieldlist(values, s_lat, s_lon, [dict(param=p, levelist=lev) for p, lev in specs])


def _same(a, b):
    assert [(f.metadata("param"), f.metadata("levelist")) for f in a] == [(f.metadata("param"), f.metadata("levelist")) for f in b]
    for fa, fb in zip(a, b):
        assert_same_values(fa.to_numpy(flatten=True), fb.to_numpy(flatten=True), str(fa.metadata("param")))
        assert np.array_equal(fa.grid_points()[0], fb.grid_points()[0])


@pytest.mark.parametrize("variant", ["nearest", "mask", "linear-recipe"])
def test_fusion_covers_gather_regrids_and_the_new_followers(F, tmp_path, variant):
    """`regrid(nearest | mask | recipe) | uv_to_ddff | q_to_r | r_to_d | cos_sin backward | clip |
    apply_mask` is ONE launch and bitwise the chain."""
    from anemoi_transform_b200.fusion import FusedRegrid
    from anemoi_transform_b200.source import FieldListSource

    s_lat, s_lon = syn.regular_latlon(2.0)
    t_lat, t_lon = syn.octahedral(24)
    data = _mixed_fields(s_lat, s_lon)
    if variant == "nearest":
        regrid = dict(method="nearest", in_grid=dict(latitudes=s_lat, longitudes=s_lon), out_grid=dict(latitudes=t_lat, longitudes=t_lon))
    elif variant == "mask":
        np.savez(tmp_path / "mask.npz", mask=np.sort(np.random.default_rng(1).choice(s_lat.size, 4000, replace=False)))
        regrid = dict(mask=str(tmp_path / "mask.npz"))
    else:
        regrid = dict(in_grid=[2.0, 2.0], out_grid="O24")
    filters = [
        F("regrid", **regrid),
        F("uv_to_ddff"),
        F("q_to_r", return_inputs="none"),
        F("r_to_d", relative_humidity="rh", temperature="t2"),
        F("cos_sin_from_rad", param="mwd").__class__.reversed(param="mwd"),
        F("clip", param="d", minimum=200.0),
        F("apply_mask", mask_param="lsm", threshold=0.5, threshold_operator=">", param=["ws", "d", "mwd"]),
    ]
    pipe = FieldListSource(dataset=data)
    for f in filters:
        pipe = pipe | f
    plan = pipe.execution_plan()
    assert len(plan) == 2 and isinstance(plan[1], FusedRegrid)
    fused = pipe.forward(None)
    assert plan[1].last_forward_was_fused
    unfused = data
    for f in filters:
        unfused = f.forward(unfused)
    _same(fused, unfused)
    params = [f.metadata("param") for f in fused]
    assert "d" in params and "mwd" in params and "lsm" not in params and "cos_mwd" not in params
    if variant != "linear-recipe":  # gathers copy: the planted -0.0 of u survives into the chain's inputs
        regridded = filters[0].forward(data)
        u = next(f for f in regridded if f.metadata("param") == "u").to_numpy(flatten=True)
        assert np.signbit(u[u == 0]).any()


def test_fused_dewpoint_backward_and_keep_inputs(F, tmp_path):
    from anemoi_transform_b200.source import FieldListSource

    s_lat, s_lon = syn.regular_latlon(2.0)
    t_lat, t_lon = syn.octahedral(24)
    d, i, p, shape = syn.bilinear_matrix(2.0, t_lat, t_lon)
    syn.save_regrid_npz(tmp_path / "m.npz", d, i, p, shape, s_lat, s_lon, t_lat, t_lon)
    rng = np.random.default_rng(3)
    t = (rng.standard_normal(s_lat.size) * 10 + 280).astype(np.float32)
    td = (t - rng.uniform(0, 15, s_lat.size)).astype(np.float32)
    data = _fieldlist([td, t, td + 1, t + 1], s_lat, s_lon, [dict(param="d", levelist=850), dict(param="t", levelist=850), dict(param="d", levelist=500), dict(param="t", levelist=500)])
    for extra in ({}, {"return_inputs": "none"}):
        filters = [F("regrid", matrix=str(tmp_path / "m.npz")), F("d_to_r", **extra)]
        pipe = FieldListSource(dataset=data) | filters[0] | filters[1]
        fused = pipe.forward(None)
        assert pipe.execution_plan()[1].last_forward_was_fused
        _same(fused, filters[1].forward(filters[0].forward(data)))
