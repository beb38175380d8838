"""SURVEY §8(f) rank 1 — rescale / convert, lnsp_to_sp, impute_nans, remove_nans, cos_sin_from_rad,
cos_sin_mean_wave_direction, r_to_d / d_to_r, sum — against the golden outputs of the imported
reference classes (tests/golden/filters_more.npz) and ports of the reference's own tests.

Exact (bitwise on non-NaN, same NaN places): rescale (mul then add, never fused), impute_nans,
sum (sequential), remove_nans (gather).  Tolerance 1e-6 of the field range for the
transcendental ones (exp / log / cos / sin / atan2 / dewpoint), angles compared circularly.
"""

import numpy as np
import pytest
from conftest import assert_close_to_range

from anemoi_transform_b200 import ekd

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


def _lev(f):
    v = f.metadata("levelist")
    return None if v is None else int(v)


def _check(result, g, name, dtype="float32", exact=(), circular=(), skip=()):
    want_order = g["order"][name]
    got_order = [[f.metadata("param"), _lev(f), str(f.to_numpy().dtype)] for f in result]
    assert got_order == want_order, (name, got_order, want_order)
    assert all(dt == dtype for _, _, dt in want_order)
    for f, want, (p, lev, _) in zip(result, g[name], want_order):
        if p in skip:
            continue
        got = f.to_numpy(flatten=True)
        if exact is True or p in exact:
            assert np.array_equal(got.astype(np.float64), want, equal_nan=True), (name, p, lev)
        else:
            assert_close_to_range(got, want, REL, f"{name} {p}@{lev}", circular=circular.get(p) if isinstance(circular, dict) else None)


PASS = ("t", "r", "lnsp", "mwd", "rad", "sst", "lsp", "cp", "sf")  # untouched fields pass through bit for bit


def test_rescale_is_bit_exact_both_ways(F, golden_filters_more):
    g = golden_filters_more
    rs = F("rescale", param="t", scale=1.8, offset=-459.67)
    fwd = rs.forward(_mixed(g))
    _check(fwd, g, "rescale_fwd", exact=True)
    _check(rs.backward(fwd), g, "rescale_bwd", exact=True)
    _check(rs.forward(_mixed(g, np.float64)), g, "rescale_fwd_f64", dtype="float64", exact=True)
    assert fwd[0].metadata("units") is None and fwd[0].metadata("param") == "t"


def test_lnsp_to_sp_and_back(F, golden_filters_more):
    g = golden_filters_more
    sp = F("lnsp_to_sp").forward(_mixed(g))
    _check(sp, g, "lnsp_to_sp", exact=PASS)
    _check(F("sp_to_lnsp").forward(sp), g, "sp_to_lnsp", exact=[p for p in PASS if p != "lnsp"])
    _check(F("lnsp_to_sp").backward(sp), g, "sp_to_lnsp", exact=[p for p in PASS if p != "lnsp"])


def test_impute_nans(F, golden_filters_more):
    g = golden_filters_more
    _check(F("impute_nans_fields", param=["sst", "cp"], value=-1.5).forward(_mixed(g)), g, "impute_sst", exact=True)
    _check(F("impute_nans", param=["sst", "cp"], value=-1.5).forward(_mixed(g)), g, "impute_sst", exact=True)
    _check(F("replace_nans", param=["sst", "cp"], value=-1.5).forward(_mixed(g)), g, "impute_sst", exact=True)


def test_cos_sin_from_rad(F, golden_filters_more):
    g = golden_filters_more
    cs = F("cos_sin_from_rad", param="rad")
    c = cs.forward(_mixed(g))
    _check(c, g, "cos_sin_from_rad", exact=PASS)
    _check(cs.backward(c), g, "rad_from_cos_sin", exact=[p for p in PASS if p != "rad"], circular={"rad": 2 * np.pi})
    bad = g["in_values"][g["order"]["in"].index(["rad", 0])].astype(np.float32).copy()
    bad[7] = 7.0
    fl = ekd.from_source("list-of-dicts", [dict(param="rad", levelist=0, values=bad.copy(), latitudes=g["lat"], longitudes=g["lon"])])
    bad[3] = 1.0  # the golden input holds a NaN there: numpy's max() would be NaN and nothing is raised
    fl_no_nan = ekd.from_source("list-of-dicts", [dict(param="rad", levelist=0, values=bad, latitudes=g["lat"], longitudes=g["lon"])])
    assert len(cs.forward(fl)) == 2
    with pytest.raises(ValueError, match="expected in radians"):
        cs.forward(fl_no_nan)


def test_cos_sin_mean_wave_direction(F, golden_filters_more):
    g = golden_filters_more
    cw = F("cos_sin_mean_wave_direction")
    c = cw.forward(_mixed(g))
    _check(c, g, "cos_sin_mwd", exact=PASS)
    _check(cw.backward(c), g, "mwd_from_cos_sin", exact=[p for p in PASS if p != "mwd"], circular={"mwd": 360.0})
    c64 = cw.forward(_mixed(g, np.float64))
    _check(c64, g, "cos_sin_mwd_f64", dtype="float64", exact=PASS)
    _check(cw.backward(c64), g, "mwd_from_cos_sin_f64", dtype="float64", exact=[p for p in PASS if p != "mwd"], circular={"mwd": 360.0})


def test_dewpoint_and_back(F, golden_filters_more):
    g = golden_filters_more
    d = F("r_to_d").forward(_mixed(g))
    # The reference writes 1e-4 into the r == 0 points of the array it is handed (dewpoint.py:63-64),
    # so its *returned input* r shows 1e-4 there; this package leaves inputs untouched (r stays 0).
    _check(d, g, "r_to_d_all", exact=[p for p in PASS if p != "r"], skip=("r",))
    names = [o[:2] for o in g["order"]["r_to_d_all"]]
    for lev in (850, 500):
        got, want = d[names.index(["r", lev])].to_numpy(flatten=True), g["r_to_d_all"][names.index(["r", lev])]
        zero = got == 0
        assert np.array_equal(got[~zero].astype(np.float64), want[~zero], equal_nan=True) and np.all(want[zero] == np.float64(np.float32(1e-4)))
    _check(F("r_to_d", return_inputs="none").forward(_mixed(g)), g, "r_to_d_none", exact=PASS)
    # backward from the reference's own d (a float32 ulp of d moves r by 2e-6 of its value, so
    # feeding our d back would test error amplification, not the kernel)
    only_dt = ekd.from_source(
        "list-of-dicts",
        [dict(param=p, levelist=lev, valid_datetime="2020-01-01T00:00:00", values=v.astype(np.float32), latitudes=g["lat"], longitudes=g["lon"])
         for (p, lev, _), v in zip(g["order"]["r_to_d_all"], g["r_to_d_all"]) if p in ("d", "t")],
    )  # fmt: skip
    _check(F("d_to_r").forward(only_dt), g, "d_to_r_all", exact=("t", "d"))
    _check(F("r_to_d").forward(_mixed(g, np.float64)), g, "r_to_d_all_f64", dtype="float64", exact=[p for p in PASS if p != "r"], skip=("r",))


def test_sum(F, golden_filters_more):
    g = golden_filters_more
    _check(F("sum", params=["lsp", "cp", "sf"], output="tp").forward(_mixed(g)), g, "sum_tp", exact=True)
    _check(F("sum", params=["lsp", "cp", "sf"], output="tp").forward(_mixed(g, np.float64)), g, "sum_tp_f64", dtype="float64", exact=True)
    with pytest.raises(ValueError, match="Missing fields"):
        F("sum", params=["lsp", "cp", "nope"], output="tp").forward(_mixed(g))
    with pytest.raises(NotImplementedError):
        F("sum", params=["lsp"], output="tp").backward(_mixed(g))


def test_remove_nans(F, golden_filters_more):
    g = golden_filters_more
    rn = F("remove_nans_fields", param="sst")
    out = rn.forward(_mixed(g))
    _check(out, g, "remove_nans_sst", exact=True)
    lat, lon = out[3].grid_points()
    assert np.array_equal(lat, g["remove_nans_lat"]) and np.array_equal(lon, g["remove_nans_lon"])
    _check(rn.forward(_mixed(g)), g, "remove_nans_sst", exact=True)  # cached mask, second call
    _check(F("remove_nans", param="sst").forward(_mixed(g)), g, "remove_nans_sst", exact=True)
    with pytest.raises(ValueError, match="not found"):
        F("remove_nans", param="nope").forward(_mixed(g))


def test_pointwise_filters_stay_device_resident_in_a_pipeline(F, golden_filters_more):
    """regrid-free chain: outputs of one filter feed the next without a host round trip, and the
    result equals applying the filters one by one from host arrays."""
    g = golden_filters_more
    chain = F("rescale", param="t", scale=1.8, offset=-459.67) | F("impute_nans", param=["sst", "cp"], value=0.0) | F("sum", params=["lsp", "cp", "sf"], output="tp")
    out = chain.forward(_mixed(g))
    x = {tuple(o): v.astype(np.float32) for o, v in zip(g["order"]["in"], g["in_values"])}
    cp = x[("cp", 0)].copy()
    cp[np.isnan(cp)] = 0.0
    want_tp = x[("lsp", 0)].copy()
    want_tp += cp
    want_tp += x[("sf", 0)]
    by = {(f.metadata("param"), _lev(f)): f for f in out}
    assert np.array_equal(by[("tp", 0)].to_numpy(flatten=True), want_tp)
    assert np.array_equal(by[("t", 850)].to_numpy(flatten=True), x[("t", 850)] * 1.8 + -459.67, equal_nan=True)


# ---- ports of the reference's own tests ------------------------------------------------------
MD = {"latitudes": [10.0, 0.0, -10.0], "longitudes": [20, 40.0], "valid_datetime": "2018-08-01T09:00:00Z"}


def _fl(specs):
    return ekd.from_source("list-of-dicts", [dict(param=p, values=np.array(v), **MD) for p, v in specs])


def _by_param(fields):
    out = {}
    for f in fields:
        out.setdefault(f.metadata("param"), []).append(f)
    return out


def test_reference_dewpoint_golden_vectors(F):
    # reference tests/field_filters/test_dewpoint.py:23-27, 47-70, 120-140
    R = [[78.13834333, 71.28598853], [99.17328572, 44.52144788], [56.49667261, 86.10495618]]
    T = [[298.42488098, 297.55574036], [278.68269348, 293.99324036], [300.61042786, 300.40144348]]
    D = [[294.34245300, 292.02214050], [278.56315613, 281.47135925], [291.19792175, 297.87370300]]
    out = _by_param(F("r_to_d").forward(_fl([("r", R), ("t", T)])))
    assert set(out) == {"r", "t", "d"} and len(out["d"]) == 1
    assert np.allclose(out["d"][0].to_numpy(), D) and out["d"][0].to_numpy().shape == (3, 2)
    assert np.array_equal(out["r"][0].to_numpy(), np.array(R)) and np.array_equal(out["t"][0].to_numpy(), np.array(T))
    back = _by_param(F("d_to_r").forward(_fl([("d", D), ("t", T)])))
    assert set(back) == {"d", "t", "r"} and np.allclose(back["r"][0].to_numpy(), R)


def test_reference_cos_sin_golden_vectors(F):
    # reference tests/field_filters/test_cos_sin_from_rad.py:23-27 and test_cos_sin_mean_wave_direction.py:23-26
    RAD = [[2.67687254, 2.59108576], [1.83746659, 1.73104875], [1.1348185, 2.23051268]]
    MWD = [[153.37349864, 148.45827835], [105.27908047, 99.18178736], [65.02031089, 127.79896253]]
    COS = [[-0.89394704, -0.85225947], [-0.26352086, -0.15956740], [0.42229696, -0.61289275]]
    SIN = [[0.44817262, 0.52311930], [0.96465370, 0.98718704], [0.90645754, 0.79016611]]
    out = _by_param(F("cos_sin_from_rad", param="RAD").forward(_fl([("RAD", RAD)])))
    assert set(out) == {"cos_RAD", "sin_RAD"}
    np.testing.assert_allclose(out["cos_RAD"][0].to_numpy(), COS, rtol=1e-6)
    np.testing.assert_allclose(out["sin_RAD"][0].to_numpy(), SIN, rtol=1e-6)
    back = _by_param(F("cos_sin_from_rad", param="some_rad", cos_param="c", sin_param="s").backward(_fl([("c", COS), ("s", SIN)])))
    assert set(back) == {"some_rad"}
    np.testing.assert_allclose(back["some_rad"][0].to_numpy(), RAD, rtol=1e-6)
    out = _by_param(F("cos_sin_mean_wave_direction").forward(_fl([("mwd", MWD)])))
    assert set(out) == {"cos_mwd", "sin_mwd"}
    assert np.allclose(out["cos_mwd"][0].to_numpy(), COS) and np.allclose(out["sin_mwd"][0].to_numpy(), SIN)
    back = _by_param(F("cos_sin_mean_wave_direction").backward(_fl([("cos_mwd", COS), ("sin_mwd", SIN)])))
    assert set(back) == {"mwd"} and np.allclose(back["mwd"][0].to_numpy(), MWD)


def test_reference_lnsp_golden_vectors(F):
    # reference tests/field_filters/test_lnsp_to_sp.py:24-26, 40-48, 70-77
    LNSP = [[1.0, 2.0], [3.0, 4.0], [5.0, 6.0]]
    out = _by_param(F("lnsp_to_sp").forward(_fl([("lnsp", LNSP)])))
    assert set(out) == {"sp"} and np.allclose(out["sp"][0].to_numpy(), np.exp(LNSP))
    back = _by_param(F("sp_to_lnsp").forward(_fl([("sp", np.exp(LNSP))])))
    assert set(back) == {"lnsp"} and np.allclose(back["lnsp"][0].to_numpy(), LNSP)
    req = F("lnsp_to_sp").patch_data_request({"param": ["sp", "t"]})
    assert req == {"param": ["t", "lnsp"]}


def test_reference_orog_to_z_golden_vectors(F):
    # reference tests/field_filters/test_orog_to_z.py:25-27, 42-78
    OROG = np.array([[243.87788459, 1892.45371246], [427.80215359, 156.92873391], [2167.93458212, 338.15794671]])
    out = _by_param(F("orog_to_z_fields").forward(_fl([("orog", OROG), ("t", OROG)])))
    assert set(out) == {"z", "t"}
    assert np.array_equal(out["z"][0].to_numpy(), OROG * 9.80665)  # float64 in, float64 out, bit-exact
    back = _by_param(F("z_to_orog").forward(_fl([("z", OROG * 9.80665)])))
    assert set(back) == {"orog"} and np.array_equal(back["orog"][0].to_numpy(), OROG * 9.80665 / 9.80665)
    f32 = OROG.astype(np.float32)
    out = _by_param(F("orog_to_z").forward(_fl([("orog", f32)])))
    assert np.array_equal(out["z"][0].to_numpy(), f32 * 9.80665) and out["z"][0].to_numpy().dtype == np.float32


def test_fieldlist_larger_than_the_device_budget_streams_in_sub_batches(F, cuda, tmp_path):
    """`regrid` on a FieldList that does not fit the HBM budget: sub-batches, outputs offloaded to
    host memory, same values as the resident path, fields still usable by the next filter."""
    from anemoi_transform_b200 import synthetic as syn

    s_lat, s_lon = syn.regular_latlon(5.0)
    t_lat, t_lon = syn.octahedral(16)
    d, i, p, shape = syn.bilinear_matrix(5.0, t_lat, t_lon)
    syn.save_regrid_npz(tmp_path / "m.npz", d, i, p, shape, s_lat, s_lon, t_lat, t_lon)
    rng = np.random.default_rng(4)
    vals = [rng.normal(280, 10, shape[1]).astype(np.float32) for _ in range(23)]
    fl = ekd.from_source("list-of-dicts", [dict(param="t", levelist=k, values=v, latitudes=s_lat, longitudes=s_lon) for k, v in enumerate(vals)])
    want = [f.to_numpy(flatten=True) for f in F("regrid", matrix=str(tmp_path / "m.npz")).forward(fl)]
    flt = F("regrid", matrix=str(tmp_path / "m.npz"))
    flt.interpolator.memory_fraction = 1e-12  # forces the minimum of 4 fields per pass
    out = flt.forward(fl)
    assert len(out) == 23
    from anemoi_transform_b200.fields import device_column_of

    assert all(device_column_of(f) is None for f in out)  # offloaded: host fields again
    for f, w, k in zip(out, want, range(23)):
        assert f.metadata("levelist") == k
        a, b = f.to_numpy(flatten=True), f.to_numpy(flatten=True)
        assert np.array_equal(a, w) and np.array_equal(b, w) and a is not b
    lat, lon = out[7].grid_points()
    assert np.array_equal(lat, t_lat) and np.array_equal(lon, t_lon)
    clipped = F("clip", param="t", minimum=275.0).forward(out)
    assert np.array_equal(clipped[22].to_numpy(flatten=True), np.clip(want[22], 275.0, None))
