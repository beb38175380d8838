"""The drop-in filters (FieldList in → FieldList out) against the golden outputs of the
imported reference filters, plus ports of the reference's own filter tests.

Tolerance of the floating-point path (BASELINE north_star): 1e-6 relative to the field's
range, float32 values; wind direction compared circularly (0° ≡ 360°); NaN / inf positions
identical.  Index-like results (ordering, metadata, dtypes, which points are masked) are exact.
"""

import numpy as np
import pytest
from conftest import assert_close_to_range, assert_same_values

from anemoi_transform_b200 import ekd
from anemoi_transform_b200 import synthetic as syn

pytestmark = pytest.mark.gpu

REL = 1e-6


@pytest.fixture(scope="module")
def F(cuda):
    from anemoi_transform_b200.filters import create_filter_by_name

    return create_filter_by_name


def _mixed(g, dtype=np.float32):
    return ekd.from_source(
        "list-of-dicts",
        [dict(param=p, levelist=lev, valid_datetime="2020-01-01T00:00:00", values=v.astype(dtype), latitudes=g["lat"], longitudes=g["lon"]) for (p, lev), v in zip(g["order"]["in"], g["in_values"])],
    )


def _check(result, g, name, dtype="float32", exact=False):
    want_order = g["order"][name]
    got_order = [[f.metadata("param"), int(f.metadata("levelist")), str(f.to_numpy().dtype)] for f in result]
    assert got_order == want_order, (got_order, want_order)
    assert all(dt == dtype for _, _, dt in want_order)
    for f, want, (p, lev, _) in zip(result, g[name], want_order):
        got = f.to_numpy(flatten=True)
        if exact:
            assert np.array_equal(got.astype(np.float64), want, equal_nan=True), (name, p, lev)
        else:
            assert_close_to_range(got, want, REL, f"{name} {p}@{lev}", circular=360.0 if p == "wdir" else None)


def test_uv_to_ddff_and_back(F, golden_filters):
    g = golden_filters
    ddff = F("uv_to_ddff").forward(_mixed(g))
    _check(ddff, g, "uv_to_ddff")
    _check(F("ddff_to_uv").forward(ddff), g, "ddff_to_uv")
    _check(F("uv_to_ddff").backward(ddff), g, "ddff_to_uv")


def test_q_to_r_and_back(F, golden_filters):
    g = golden_filters
    r_all = F("q_to_r").forward(_mixed(g))
    _check(r_all, g, "q_to_r_all")
    _check(F("q_to_r", return_inputs="none").forward(_mixed(g)), g, "q_to_r_none")
    only_rt = ekd.SimpleFieldList([f for f in r_all if f.metadata("param") in ("r", "t")])
    _check(F("r_to_q").forward(only_rt), g, "r_to_q_all")


def test_float64_fields_stay_float64(F, golden_filters):
    g = golden_filters
    _check(F("uv_to_ddff").forward(_mixed(g, np.float64)), g, "uv_to_ddff_f64", dtype="float64")
    _check(F("q_to_r").forward(_mixed(g, np.float64)), g, "q_to_r_all_f64", dtype="float64")


def test_clip(F, golden_filters):
    g = golden_filters
    _check(F("clip", param="t", minimum=250.0, maximum=280.0).forward(_mixed(g)), g, "clip_t_both", exact=True)
    _check(F("clipper", param="q", minimum=0.005).forward(_mixed(g)), g, "clip_q_min", exact=True)
    _check(F("clip_fields", param="u", maximum=-1.5).forward(_mixed(g)), g, "clip_u_max", exact=True)
    x = np.array([np.nan, 1.0, 5.0, -np.inf, np.inf], dtype=np.float32)
    fl = ekd.from_source("list-of-dicts", [dict(param="tp", levelist=0, values=x, latitudes=np.zeros(5), longitudes=np.arange(5.0))])
    out = F("clip", param="tp", minimum=2.0, maximum=4.0).forward(fl)[0].to_numpy()
    assert_same_values(out, np.clip(x, 2.0, 4.0))  # NaN passes through np.clip


def test_apply_mask_from_param(F, golden_filters):
    g = golden_filters
    _check(F("apply_mask", mask_param="lsm", mask_value=0).forward(_mixed(g)), g, "mask_param_value", exact=True)
    _check(F("mask", mask_param="lsm", threshold=0.5, threshold_operator="<=", return_mask=True, param=["t", "q"], rename="land").forward(_mixed(g)), g, "mask_param_thr_keep", exact=True)
    _check(F("apply_mask_fields", mask_param="z", threshold=5e4, threshold_operator="gt", param="u").forward(_mixed(g)), g, "mask_param_ne", exact=True)


# ---- ports of the reference's own tests ------------------------------------------------------
MD = {"latitudes": [10.0, 0.0, -10.0], "longitudes": [20, 40.0], "valid_datetime": "2018-08-01T09:00:00Z"}
U = {500: [[-3.26786804, -2.90458679], [-4.28153992, -10.75224304], [-6.29130554, -4.17704773]], 850: [[-6.72481718, -0.34174164], [-7.14725727, -2.02047454], [-4.93597360, -0.00018431]]}
V = {500: [[6.51824951, 4.7321167], [1.16961670, 1.73797607], [-2.93096924, 3.2399292]], 850: [[5.4374572, -0.00679462], [2.23226754, 6.78457592], [-1.79188286, -0.0093771]]}
WS = {500: [[7.29153881, 5.55243666], [4.43842171, 10.89179926], [6.94054076, 5.28629066]], 850: [[8.64806955, 0.34180918], [7.48774364, 7.07903862], [5.25115983, 0.00937891]]}
WDIR = {500: [[153.37349864, 148.45827835], [105.27908047, 99.18178736], [65.02031089, 127.79896253]], 850: [[128.95781905, 88.86097611], [107.34489249, 163.41625261], [70.04782648, 1.12603196]]}
T = {850: [[293.32301331, 284.21559143], [260.53981018, 291.18824768], [279.88941956, 248.87574768]], 1000: [[291.22831726, 289.85136414], [271.29277039, 301.67362976], [287.53691101, 250.15409851]]}
Q = {850: [[0.00657578, 0.00769957], [0.00147607, 0.01088967], [0.00505508, 0.00044559]], 1000: [[0.01075057, 0.01080445], [0.00226020, 0.01525551], [0.00914679, 0.00047560]]}
R = {850: [[37.91091442, 79.51638317], [95.61794567, 71.53396130], [70.03982067, 89.69021130]], 1000: [[82.88058853, 90.86496353], [68.26144791, 62.40207291], [89.31613541, 99.25949478]]}


def _src(specs):
    from anemoi_transform_b200.source import FieldListSource

    return FieldListSource(dataset=ekd.from_source("list-of-dicts", [dict(param=p, levelist=lev, values=np.array(v), **MD) for p, lev, v in specs]))


def _by_param(pipeline):
    out = {}
    for f in pipeline:
        out.setdefault(f.metadata("param"), []).append(f)
    return out


def test_reference_uv_golden_vectors_and_round_trip(F):
    # reference tests/field_filters/test_uv_to_ddff.py:67-132
    src = _src([("u", 500, U[500]), ("v", 500, V[500]), ("u", 850, U[850]), ("v", 850, V[850])])
    out = _by_param(src | F("uv_to_ddff"))
    assert set(out) == {"ws", "wdir"} and len(out["ws"]) == 2 and len(out["wdir"]) == 2
    for i, lev in enumerate([500, 850]):
        assert np.allclose(out["ws"][i].to_numpy(), WS[lev]) and np.allclose(out["wdir"][i].to_numpy(), WDIR[lev])
        assert out["ws"][i].metadata("levelist") == lev and out["ws"][i].to_numpy().shape == (3, 2)
    back = _by_param(src | F("uv_to_ddff") | F("ddff_to_uv"))
    assert set(back) == {"u", "v"}
    for i, lev in enumerate([500, 850]):
        assert np.allclose(back["u"][i].to_numpy(), U[lev]) and np.allclose(back["v"][i].to_numpy(), V[lev])
    src2 = _src([("ws", 500, WS[500]), ("wdir", 500, WDIR[500]), ("ws", 850, WS[850]), ("wdir", 850, WDIR[850])])
    out2 = _by_param(src2 | F("ddff_to_uv"))
    for i, lev in enumerate([500, 850]):
        assert np.allclose(out2["u"][i].to_numpy(), U[lev]) and np.allclose(out2["v"][i].to_numpy(), V[lev])


def test_reference_humidity_golden_vectors_and_round_trip(F):
    # reference tests/field_filters/test_pressure_level_humidity.py:65-115
    src = _src([("q", 850, Q[850]), ("t", 850, T[850]), ("q", 1000, Q[1000]), ("t", 1000, T[1000])])
    out = _by_param(src | F("q_to_r"))
    assert set(out) == {"q", "t", "r"}
    for f in out["r"]:
        assert np.allclose(f.to_numpy(), R[f.metadata("levelist")])
    for p, ref in (("q", Q), ("t", T)):
        for f in out[p]:
            assert np.array_equal(f.to_numpy(), np.array(ref[f.metadata("levelist")]))
    rt = ekd.SimpleFieldList([f for f in (src | F("q_to_r")) if f.metadata("param") in ("r", "t")])
    back = _by_param(F("r_to_q").forward(rt))
    assert set(back) == {"q", "t", "r"}
    for f in back["q"]:
        assert np.allclose(f.to_numpy(), Q[f.metadata("levelist")])


LSM = np.array([[1, 0], [1, 1], [0, 0]])
DATA = {"sd": np.array([[1.0, 2.0], [3.0, 4.0], [5.0, 6.0]]), "lsm": LSM.astype(float), "2t": np.array([[7.0, 8.0], [9.0, 0.0], [9.0, 8.0]])}


def _mask_src():
    from anemoi_transform_b200.source import FieldListSource

    return FieldListSource(dataset=ekd.from_source("list-of-dicts", [dict(param=p, values=v.copy(), **MD) for p, v in DATA.items()]))


def _expect_masked(field, param, expected_mask):
    expected = DATA[param].flatten().copy()
    expected[expected_mask] = np.nan
    assert np.array_equal(field.to_numpy(flatten=True), expected, equal_nan=True)


def test_reference_apply_mask_from_field(F):
    # reference tests/field_filters/test_apply_mask_from_field.py:40-146
    out = _by_param(_mask_src() | F("apply_mask", mask_param="lsm", mask_value=0))
    assert "lsm" not in out
    for p in ("sd", "2t"):
        _expect_masked(out[p][0], p, LSM.flatten() == 0)
    out = _by_param(_mask_src() | F("apply_mask", mask_param="lsm", threshold=0.5, threshold_operator="<", return_mask=False))
    for p in ("sd", "2t"):
        _expect_masked(out[p][0], p, LSM.flatten() < 0.5)
    out = _by_param(_mask_src() | F("apply_mask", mask_param="lsm", mask_value=0, param="sd", return_mask=False))
    _expect_masked(out["sd"][0], "sd", LSM.flatten() == 0)
    assert np.array_equal(out["2t"][0].to_numpy(flatten=True), DATA["2t"].flatten())
    with pytest.raises(ValueError, match="not found in input data"):
        list(_mask_src() | F("apply_mask", mask_param="nonexistent", mask_value=0))
    out = _by_param(_mask_src() | F("apply_mask", mask_param="lsm", mask_value=0, return_mask=True, param=["sd", "2t"]))
    assert len(out["lsm"]) == 1 and np.array_equal(out["lsm"][0].to_numpy(flatten=True), LSM.flatten())
    for p in ("sd", "2t"):
        _expect_masked(out[p][0], p, LSM.flatten() == 0)


def test_apply_mask_from_npy_file(F, tmp_path):
    # reference tests/field_filters/test_apply_mask.py (file mask; .npy branch apply_mask.py:154-155)
    np.save(tmp_path / "mask.npy", LSM.flatten())
    out = _by_param(_mask_src() | F("apply_mask", path=str(tmp_path / "mask.npy"), mask_value=0, rename="masked"))
    assert set(out) == {"sd_masked", "lsm_masked", "2t_masked"}
    _expect_masked(out["sd_masked"][0], "sd", LSM.flatten() == 0)
    out = _by_param(_mask_src() | F("apply_mask", path=str(tmp_path / "mask.npy"), threshold=0, threshold_operator="ne", param="2t"))
    _expect_masked(out["2t"][0], "2t", LSM.flatten() != 0)
    assert np.array_equal(out["sd"][0].to_numpy(), DATA["sd"])


# ---- regrid filter ---------------------------------------------------------------------------
def _regrid_src(g, key="fields32"):
    names = ["t", "u", "v", "q", "t", "z"]
    return ekd.from_source(
        "list-of-dicts",
        [dict(param=names[i], levelist=500 + i, valid_datetime="2020-01-01T00:00:00", values=v, latitudes=g["s_lat"], longitudes=g["s_lon"]) for i, v in enumerate(g[key])],
    )


def test_regrid_matrix_filter_matches_reference_filter(F, golden_regrid, tmp_path):
    g = golden_regrid
    syn.save_regrid_npz(tmp_path / "m32.npz", g["m32_data"], g["m32_indices"], g["m32_indptr"], g["m32_shape"], g["s_lat"], g["s_lon"], g["t_lat"], g["t_lon"])
    syn.save_regrid_npz(tmp_path / "m64.npz", g["m64_data"], g["m64_indices"], g["m64_indptr"], g["m64_shape"], g["s_lat"], g["s_lon"], g["irr_lat"], g["irr_lon"])
    out = F("regrid", matrix=str(tmp_path / "m32.npz")).forward(_regrid_src(g))
    assert [f.metadata("param") for f in out] == ["t", "u", "v", "q", "t", "z"]
    assert [f.metadata("levelist") for f in out] == [500, 501, 502, 503, 504, 505]
    assert_same_values(np.stack([f.to_numpy(flatten=True) for f in out]), g["y_m32_f32"], "regrid m32")
    lat, lon = out[0].grid_points()
    assert np.array_equal(lat, g["y_m32_lat"]) and np.array_equal(lon, g["y_m32_lon"]) and out[0].shape == (1600,)
    assert np.array_equal(out[2].metadata().geography.latitudes(), g["t_lat"])
    out = F("regrid", matrix=str(tmp_path / "m64.npz")).forward(_regrid_src(g))
    assert_same_values(np.stack([f.to_numpy(flatten=True) for f in out]), g["y_m64_f32"], "regrid m64 (float64 matrix → float64)")
    out = F("regrid", matrix=str(tmp_path / "m32.npz")).forward(_regrid_src(g, "fields64"))
    assert_same_values(np.stack([f.to_numpy(flatten=True) for f in out]), g["y_m32_f64"], "regrid float64 fields")
    assert len(F("regrid", matrix=str(tmp_path / "m32.npz")).forward(ekd.SimpleFieldList([]))) == 0


def test_regrid_nearest_and_mask_filters(F, golden_regrid, tmp_path):
    g = golden_regrid
    out = F("regrid", method="nearest", in_grid=dict(latitudes=g["s_lat"], longitudes=g["s_lon"]), out_grid=dict(latitudes=g["t_lat"], longitudes=g["t_lon"])).forward(_regrid_src(g))
    got = np.stack([f.to_numpy(flatten=True) for f in out])
    # a 5° lat-lon source has targets with two sources at exactly the same float64 distance;
    # there cKDTree's pick is its traversal order and ours is the lowest index (tie-flagged)
    from anemoi_transform_b200 import spatial

    idx, _, ties = spatial.nearest_grid_points(g["s_lat"], g["s_lon"], g["t_lat"], g["t_lon"], _return_ties=True)
    untied = ties == 0
    assert untied.sum() > 0.99 * untied.size
    assert_same_values(got[:, untied], g["y_nearest_f32"][:, untied], "nearest (untied targets)")
    assert_same_values(got, g["fields32"][:, idx], "nearest = gather of our own indices")
    assert np.array_equal(out[0].grid_points()[0], g["y_nearest_lat"])
    np.savez(tmp_path / "mask.npz", mask=g["mask_idx"])
    out = F("regrid", mask=str(tmp_path / "mask.npz")).forward(_regrid_src(g))
    assert_same_values(np.stack([f.to_numpy(flatten=True) for f in out]), g["y_mask_f32"], "mask")
    assert np.array_equal(out[0].grid_points()[0], g["y_mask_lat"]) and np.array_equal(out[0].grid_points()[1], g["y_mask_lon"])
    boolean = np.zeros(g["s_lat"].size, dtype=bool)
    boolean[g["mask_idx"]] = True
    np.savez(tmp_path / "maskb.npz", mask=boolean)
    out = F("regrid", mask=str(tmp_path / "maskb.npz")).forward(_regrid_src(g))
    assert_same_values(np.stack([f.to_numpy(flatten=True) for f in out]), g["y_mask_f32"], "boolean mask")
    np.savez(tmp_path / "maskbad.npz", mask=np.array([0, g["s_lat"].size]))
    with pytest.raises(IndexError):
        F("regrid", mask=str(tmp_path / "maskbad.npz")).forward(_regrid_src(g))


def test_pipeline_stays_on_device_and_matches_oracle_chain(F, golden_regrid, tmp_path):
    """regrid | uv_to_ddff | q_to_r | clip | apply_mask on device-resident fields against the
    same chain of oracle steps on the host."""
    from anemoi_transform_b200.fields import device_column_of
    from anemoi_transform_b200.source import FieldListSource
    from oracle import pointwise as pw

    g = golden_regrid
    syn.save_regrid_npz(tmp_path / "m32.npz", g["m32_data"], g["m32_indices"], g["m32_indptr"], g["m32_shape"], g["s_lat"], g["s_lon"], g["t_lat"], g["t_lon"])
    n_s = g["s_lat"].size
    vals = {"u": syn.synthetic_field("u", n_s, 1), "v": syn.synthetic_field("v", n_s, 2), "q": syn.synthetic_field("q", n_s, 3), "t": syn.synthetic_field("t", n_s, 4), "lsm": syn.synthetic_field("lsm", n_s, 5)}
    src = FieldListSource(dataset=ekd.from_source("list-of-dicts", [dict(param=p, levelist=850, values=v, latitudes=g["s_lat"], longitudes=g["s_lon"]) for p, v in vals.items()]))
    pipe = src | F("regrid", matrix=str(tmp_path / "m32.npz")) | F("uv_to_ddff") | F("q_to_r") | F("clip", param="r", minimum=0.0, maximum=100.0) | F("apply_mask", mask_param="lsm", threshold=0.5, threshold_operator=">")
    out = pipe.forward(None)
    assert all(device_column_of(f) is not None for f in out)  # nothing went back to the host
    got = {f.metadata("param"): f.to_numpy(flatten=True) for f in out}
    assert list(got) == ["ws", "wdir", "q", "t", "r"]

    from scipy.sparse import csr_array

    m = csr_array((g["m32_data"], g["m32_indices"], g["m32_indptr"]), shape=tuple(g["m32_shape"]))
    reg = {p: m @ v for p, v in vals.items()}
    ws, wdir = pw.xy_to_polar(reg["u"], reg["v"])
    r = np.clip(pw.relative_humidity_from_specific_humidity(reg["t"], reg["q"], 85000.0), 0.0, 100.0)
    mask = reg["lsm"] > 0.5
    want = {"ws": ws, "wdir": wdir, "q": reg["q"], "t": reg["t"], "r": r}
    for p, w in want.items():
        assert_close_to_range(got[p], pw.apply_mask(w, mask), REL, p, circular=360.0 if p == "wdir" else None)
    assert 0 < mask.sum() < mask.size


def _fusion_inputs(g, tmp_path):
    from anemoi_transform_b200.source import FieldListSource

    syn.save_regrid_npz(tmp_path / "m32.npz", g["m32_data"], g["m32_indices"], g["m32_indptr"], g["m32_shape"], g["s_lat"], g["s_lon"], g["t_lat"], g["t_lon"])
    n_s = g["s_lat"].size
    specs = [("t", 850), ("u", 850), ("z", 500), ("v", 850), ("u", 500), ("q", 850), ("v", 500), ("t", 500), ("q", 500), ("lsm", 0), ("u", 300), ("v", 300)]
    fields = [dict(param=p, levelist=lev, values=syn.synthetic_field(p, n_s, 40 + i, 0.002 if i % 3 == 0 else 0.0), latitudes=g["s_lat"], longitudes=g["s_lon"]) for i, (p, lev) in enumerate(specs)]
    return FieldListSource(dataset=ekd.from_source("list-of-dicts", fields)), str(tmp_path / "m32.npz")


def _run_unfused(filters, data):
    for f in filters:
        data = f.forward(data)
    return data


def _same_fieldlists(a, b):
    assert [(f.metadata("param"), f.metadata("levelist")) for f in a] == [(f.metadata("param"), f.metadata("levelist")) for f in b]
    for fa, fb in zip(a, b):
        assert_same_values(fa.to_numpy(flatten=True), fb.to_numpy(flatten=True), str(fa.metadata("param")))
        assert np.array_equal(fa.grid_points()[0], fb.grid_points()[0])


def test_pipeline_fusion_one_launch_same_result(F, golden_regrid, tmp_path):
    """`regrid | uv_to_ddff | q_to_r | clip | clip | apply_mask` runs as ONE at_spmm_fused
    launch and gives bitwise what the six filters give one after the other."""
    from anemoi_transform_b200.fusion import FusedRegrid

    src, matrix = _fusion_inputs(golden_regrid, tmp_path)
    filters = [F("regrid", matrix=matrix), F("uv_to_ddff"), F("q_to_r"), F("clip", param="r", minimum=0.0, maximum=100.0), F("clip", param="ws", maximum=25.0), F("apply_mask", mask_param="lsm", threshold=0.5, threshold_operator=">", rename="land", param=["ws", "r", "z"])]
    pipe = src
    for f in filters:
        pipe = pipe | f
    plan = pipe.execution_plan()
    assert len(plan) == 2 and isinstance(plan[1], FusedRegrid)
    fused = pipe.forward(None)
    assert plan[1].last_forward_was_fused
    _same_fieldlists(fused, _run_unfused(filters, src.forward(None)))
    params = [f.metadata("param") for f in fused]
    assert "lsm" not in params and params.count("ws_land") == 3 and params.count("r_land") == 2 and params[0] == "z_land"
    # return_inputs="none", reversed filters and a file mask fuse as well
    np.save(tmp_path / "mask.npy", (np.random.default_rng(1).uniform(size=golden_regrid["t_lat"].size) < 0.2).astype(np.float32))
    filters = [F("regrid", matrix=matrix), F("q_to_r", return_inputs="none"), F("uv_to_ddff"), F("apply_mask", path=str(tmp_path / "mask.npy"), mask_value=1)]
    pipe = src | filters[0] | filters[1] | filters[2] | filters[3]
    fused = pipe.forward(None)
    assert pipe.execution_plan()[1].last_forward_was_fused
    _same_fieldlists(fused, _run_unfused(filters, src.forward(None)))


def test_pipeline_fusion_of_unary_filters(F, golden_regrid, tmp_path):
    """`regrid | rescale | orog_to_z | impute_nans | uv_to_ddff | clip | apply_mask` — the one-field
    filters of SURVEY §8(f) rank 1 fuse into the same single launch, bitwise equal to the chain."""
    from anemoi_transform_b200.fusion import FusedRegrid

    src, matrix = _fusion_inputs(golden_regrid, tmp_path)
    filters = [
        F("regrid", matrix=matrix),
        F("rescale", param="t", scale=1.8, offset=-459.67),
        F("z_to_orog", orography="h"),
        F("impute_nans", param=["q", "z"], value=0.0),
        F("uv_to_ddff"),
        F("clip", param="t", maximum=60.0),
        F("apply_mask", mask_param="lsm", threshold=0.5, threshold_operator=">", param=["h", "ws"]),
    ]
    pipe = src
    for f in filters:
        pipe = pipe | f
    plan = pipe.execution_plan()
    assert len(plan) == 2 and isinstance(plan[1], FusedRegrid)
    fused = pipe.forward(None)
    assert plan[1].last_forward_was_fused
    _same_fieldlists(fused, _run_unfused(filters, src.forward(None)))
    params = [f.metadata("param") for f in fused]
    assert "z" not in params and "h" in params and "ws" in params and "lsm" not in params
    # the reversed forms fuse too; a conversion of a converted field does not, and still agrees
    filters = [F("regrid", matrix=matrix), F("rescale", param="t", scale=1.8, offset=-459.67).__class__.reversed(param="t", scale=1.8, offset=-459.67), F("sp_to_lnsp", surface_pressure="z")]
    pipe = src | filters[0] | filters[1] | filters[2]
    fused = pipe.forward(None)
    assert pipe.execution_plan()[1].last_forward_was_fused
    _same_fieldlists(fused, _run_unfused(filters, src.forward(None)))
    filters = [F("regrid", matrix=matrix), F("rescale", param="t", scale=2.0, offset=1.0), F("rescale", param="t", scale=0.5, offset=0.0)]
    pipe = src | filters[0] | filters[1] | filters[2]
    out = pipe.forward(None)
    assert not pipe.execution_plan()[1].last_forward_was_fused
    _same_fieldlists(out, _run_unfused(filters, src.forward(None)))


def test_pipeline_fusion_falls_back_when_the_epilogue_cannot_express_it(F, golden_regrid, tmp_path):
    src, matrix = _fusion_inputs(golden_regrid, tmp_path)
    # a clip BEFORE the conversion, and two clips of the same field: not expressible, still correct
    for follow in ([F("clip", param="u", minimum=-5.0), F("uv_to_ddff")], [F("clip", param="t", minimum=250.0), F("clip", param="t", maximum=300.0)]):
        filters = [F("regrid", matrix=matrix)] + follow
        pipe = src | filters[0] | filters[1] | filters[2]
        out = pipe.forward(None)
        assert not pipe.execution_plan()[1].last_forward_was_fused
        _same_fieldlists(out, _run_unfused(filters, src.forward(None)))
    # errors of the followers surface unchanged
    with pytest.raises(ValueError, match="not found in input data"):
        (src | F("regrid", matrix=matrix) | F("apply_mask", mask_param="nope", mask_value=0)).forward(None)
    with pytest.raises(ValueError, match="Missing component"):
        only_u = ekd.SimpleFieldList([f for f in src.forward(None) if f.metadata("param") != "v"])
        (F("regrid", matrix=matrix) | F("uv_to_ddff")).forward(only_u)


def test_fused_epilogue_equals_unfused_chain(cuda, golden_regrid):
    """at_spmm_fused (regrid + uv_to_ddff + q_to_r(all) + clip + mask in one kernel) gives
    bitwise the same numbers as at_spmm followed by at_pointwise."""
    from anemoi_transform_b200 import _cabi
    from anemoi_transform_b200.device import CsrMatrix, DeviceBatch, Epilogue

    g = golden_regrid
    csr = CsrMatrix(g["m32_data"], g["m32_indices"], g["m32_indptr"], tuple(g["m32_shape"]))
    n_s = g["s_lat"].size
    cols = [syn.synthetic_field(p, n_s, 10 + i, 0.002) for i, p in enumerate(["t", "z", "t", "z", "u", "v", "u", "v", "q", "t", "q", "t"])]
    x = DeviceBatch.from_host_fields(cols).data
    CL, CH, MK = _cabi.COL_CLIP_LO, _cabi.COL_CLIP_HI, _cabi.COL_MASK
    segments = [(_cabi.EPI_PLAIN, 0, 4, 0), (_cabi.EPI_UV2DDFF, 4, 4, 4), (_cabi.EPI_QT2QTR, 8, 4, 8)]
    out_cols = [(270.0, 290.0, 0, CL | CH), (0, 0, 0, MK), (0, 0, 0, 0), (0, 5e4, 0, CH | MK)]
    out_cols += [(0, 0, 0, MK), (0, 0, 0, 0), (2.0, 0, 0, CL), (0, 0, 0, 0)]
    out_cols += [(0, 0, 0, 0), (0, 0, 0, 0), (0, 100.0, 85000.0, CL | CH | MK), (0, 0, 0, 0), (0, 0, 0, 0), (0, 100.0, 50000.0, CL | CH)]
    out_cols += [(0, 0, 0, 0)] * 2
    epi = Epilogue(segments, out_cols)
    row_mask = cuda.from_numpy((np.random.default_rng(0).uniform(size=csr.shape[0]) < 0.3).astype(np.uint8)).cuda()
    fused = epi.apply_fused(csr, x, row_mask=row_mask)
    unfused = epi.apply(csr.apply(x), row_mask=row_mask)
    assert_same_values(fused[:, :14].cpu().numpy(), unfused[:, :14].cpu().numpy(), "fused vs unfused")
    y = csr.apply(x).cpu().numpy()
    f = fused.cpu().numpy()
    rm = row_mask.cpu().numpy().astype(bool)
    assert_same_values(f[:, 0], np.clip(y[:, 0], np.float32(270.0), np.float32(290.0)), "clip column")
    assert np.isnan(f[rm, 1]).all() and np.array_equal(f[~rm, 1], y[~rm, 1], equal_nan=True)
    assert np.array_equal(f[:, 8], y[:, 8], equal_nan=True) and np.array_equal(f[:, 9], y[:, 9], equal_nan=True)  # q, t returned
