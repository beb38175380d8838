"""GRIB simple packing decoded on the device (csrc/grib.cu) against the numpy oracle, and the
regrid filter fed with GRIB-backed fields against the same filter fed with the host-decoded
values.  Bit-exact throughout (float64 decode: one multiply, one add, one multiply)."""

import numpy as np
import pytest
from conftest import assert_same_values
from grib_fields import GribMessageField
from scipy.sparse import csr_array

from anemoi_transform_b200 import synthetic as syn
from oracle import grib as ogrib

pytestmark = pytest.mark.gpu


def _messages(n_fields, n_points, widths, decimals, editions, seed=0):
    rng = np.random.default_rng(seed)
    out = []
    for k in range(n_fields):
        nb, d, ed = widths[k % len(widths)], decimals[k % len(decimals)], editions[k % len(editions)]
        v = rng.normal(280.0, 15.0, n_points) if k % 2 else rng.uniform(-40.0, 40.0, n_points)
        if nb == 0:
            v = np.full(n_points, 2.5 * k)
        out.append((ogrib.encode_grib2 if ed == 2 else ogrib.encode_grib1)(v, max(nb, 1), d))
    return out


def _fields(messages, n_points, lat=None, lon=None):
    return [GribMessageField(m, n_points, dict(param="t", levelist=850, step=k), latitudes=lat, longitudes=lon) for k, m in enumerate(messages)]


@pytest.mark.parametrize("n_points", [1, 7, 8, 255, 256, 1001, 65160])
def test_every_width_decodes_like_the_oracle(cuda, n_points):
    from anemoi_transform_b200 import grib

    widths = [0, 1, 3, 7, 8, 11, 12, 13, 16, 20, 24, 28, 31, 32]
    msgs = _messages(len(widths) * 2, n_points, widths, [0], [2, 2, 1, 1], seed=n_points)
    want = np.stack([ogrib.decode(m, n_points=n_points) for m in msgs])
    packed = grib.packed_of(_fields(msgs, n_points))
    assert packed is not None
    batch = grib.upload(packed)
    assert batch.data.dtype == cuda.float64 and batch.n_fields == len(msgs) and batch.n_points == n_points
    assert_same_values(batch.data[:, : len(msgs)].cpu().numpy().T, want, "float64 decode")
    as_f32 = grib.upload(packed, np.float32)
    assert_same_values(as_f32.data[:, : len(msgs)].cpu().numpy().T, want.astype(np.float32), "float32 decode")


def test_decimal_scale_factors_and_many_fields(cuda):
    from anemoi_transform_b200 import grib

    n_points = 40320
    msgs = _messages(131, n_points, [16, 12, 24, 9], [0, 2, -1, 5, -3], [2, 1], seed=3)
    want = np.stack([ogrib.decode(m, n_points=n_points) for m in msgs])
    got = grib.upload(grib.packed_of(_fields(msgs, n_points))).data[:, :131].cpu().numpy().T
    assert_same_values(got, want, "decimal scale")


def test_long_edition1_message(cuda):
    from anemoi_transform_b200 import grib

    n_points = 4_300_000
    v = np.random.default_rng(1).normal(101325.0, 900.0, n_points)
    msg = ogrib.encode_grib1(v, 16, 0)
    assert len(msg) > 0x800000
    got = grib.upload(grib.packed_of(_fields([msg], n_points))).data[:, 0].cpu().numpy()
    assert_same_values(got, ogrib.decode(msg, n_points=n_points), "long GRIB1")


def test_raw_unpack_entry_point_validates(cuda):
    """at_grib_unpack on packed values already in device memory; bad arguments are refused."""
    import ctypes

    from anemoi_transform_b200 import _cabi, grib
    from anemoi_transform_b200.device import _ptr, stream_ptr

    n_points = 999
    msgs = _messages(5, n_points, [16, 10], [0], [2])
    infos = (_cabi.GribInfo * 5)(*[grib.scan(m) for m in msgs])
    blob, offs = bytearray(), []
    for m, i in zip(msgs, infos):
        while len(blob) % 256:
            blob.append(0)
        offs.append(len(blob))
        blob += m[i.data_offset : i.data_offset + i.data_length]
    d_blob = cuda.frombuffer(blob, dtype=cuda.uint8).cuda()
    out = cuda.empty((n_points, 8), dtype=cuda.float64, device="cuda")
    offsets = (ctypes.c_int64 * 5)(*offs)
    _cabi.call("at_grib_unpack", _ptr(d_blob), offsets, infos, 5, n_points, _cabi.AT_F64, _ptr(out), 8, stream_ptr())
    want = np.stack([ogrib.decode(m) for m in msgs])
    assert_same_values(out[:, :5].cpu().numpy().T, want, "at_grib_unpack")
    with pytest.raises(ValueError, match="values"):  # the batch has another number of points
        _cabi.call("at_grib_unpack", _ptr(d_blob), offsets, infos, 5, n_points + 1, _cabi.AT_F64, _ptr(out), 8, stream_ptr())
    with pytest.raises(ValueError, match="leading dimension"):
        _cabi.call("at_grib_unpack", _ptr(d_blob), offsets, infos, 5, n_points, _cabi.AT_F64, _ptr(out), 4, stream_ptr())
    # packed values alone are not enough for a message with a bitmap: that goes through the engine
    bm = np.random.default_rng(0).uniform(size=n_points) > 0.5
    with_bitmap = ogrib.encode_grib2(np.arange(float(bm.sum())), 16, 0, bm)
    one = (_cabi.GribInfo * 1)(grib.scan(with_bitmap))
    with pytest.raises(ValueError, match="bitmap"):
        _cabi.call("at_grib_unpack", _ptr(d_blob), offsets, one, 1, n_points, _cabi.AT_F64, _ptr(out), 8, stream_ptr())


@pytest.mark.parametrize("n_points", [5, 255, 256, 257, 1001, 40320])
def test_bitmaps_decode_to_nan_where_the_bit_is_clear(cuda, n_points):
    """Fields with a bitmap (editions 1 and 2, several widths, all-present / all-missing /
    random / striped masks) next to fields without one: values at the set bits, NaN elsewhere."""
    from anemoi_transform_b200 import grib

    rng = np.random.default_rng(n_points)
    masks = [rng.uniform(size=n_points) > 0.3, np.ones(n_points, bool), np.zeros(n_points, bool), np.arange(n_points) % 3 == 0, rng.uniform(size=n_points) > 0.97]
    masks[2][n_points // 2] = True  # (an encoder needs at least one value)
    msgs = []
    for k, bm in enumerate(masks * 2):
        nb = (16, 12, 24, 7, 32, 9, 8, 16, 11, 20)[k]
        enc = ogrib.encode_grib2 if k % 2 == 0 else ogrib.encode_grib1
        msgs.append(enc(rng.normal(285.0, 5.0, int(bm.sum())), nb, 0 if k % 3 else 1, bm))
        msgs.append(enc(rng.normal(0.0, 8.0, n_points), nb, 0))  # and a field without a bitmap in between
    want = np.stack([ogrib.decode(m, n_points=n_points) for m in msgs])
    assert np.isnan(want).any()
    packed = grib.packed_of(_fields(msgs, n_points))
    assert packed is not None and packed.n_fields == len(msgs)
    got = grib.upload(packed).data[:, : len(msgs)].cpu().numpy().T
    assert_same_values(got, want, "bitmap decode float64")
    got32 = grib.upload(packed, np.float32).data[:, : len(msgs)].cpu().numpy().T
    assert_same_values(got32, want.astype(np.float32), "bitmap decode float32")


@pytest.mark.parametrize("mdtype", [np.float32, np.float64])
def test_regrid_of_grib_fields_equals_regrid_of_decoded_fields(cuda, tmp_path, mdtype):
    """FieldList of GRIB-backed fields -> RegridFilter.forward -> to_numpy: the packed messages
    are decoded on the device (no host decode at all) and the result is scipy's `matrix @
    to_numpy()` bit for bit, streamed and inside a pipeline."""
    from anemoi_transform_b200 import ekd
    from anemoi_transform_b200.filters import create_filter_by_name as F

    t_lat, t_lon = syn.octahedral(48)
    d, i, p, shape = syn.bilinear_matrix(2.0, t_lat, t_lon)
    d = d.astype(mdtype)
    s_lat, s_lon = syn.regular_latlon(2.0)
    path = str(tmp_path / "m.npz")
    syn.save_regrid_npz(path, d, i, p, shape, s_lat, s_lon, t_lat, t_lon)
    m = csr_array((d, i, p), shape=shape)
    n_src = shape[1]
    msgs = _messages(37, n_src, [16, 12, 24], [0, 1], [2, 1], seed=11)
    fields = _fields(msgs, n_src, s_lat, s_lon)
    want = [m @ ogrib.decode(msg, n_points=n_src) for msg in msgs]

    out = F("regrid", matrix=path).forward(ekd.SimpleFieldList(fields))
    assert len(out) == 37
    for k, f in enumerate(out):
        got = f.to_numpy()
        assert got.dtype == np.float64
        assert_same_values(got, want[k], f"streamed field {k}")
        assert np.array_equal(f.grid_points()[0], t_lat) and f.metadata("step") == k
    assert sum(f.decodes for f in fields) == 0  # nothing was decoded on the host

    # as the first filter of a pipeline the results stay resident and feed the next filter
    pipe = F("regrid", matrix=path) | F("rescale", param="t", scale=2.0, offset=1.0)
    out = pipe.forward(ekd.SimpleFieldList(fields))
    for k, f in enumerate(out):
        assert_same_values(f.to_numpy(), want[k] * 2.0 + 1.0, f"pipeline field {k}")
    assert sum(f.decodes for f in fields) == 0

    # nearest-neighbour regrid: a gather of the decoded values
    near = F("regrid", method="nearest", in_grid=dict(latitudes=s_lat, longitudes=s_lon), out_grid=dict(latitudes=t_lat, longitudes=t_lon))
    ref = near.forward(ekd.SimpleFieldList([ekd.ArrayField(ogrib.decode(mm, n_points=n_src), dict(param="t", step=k), latitudes=s_lat, longitudes=s_lon) for k, mm in enumerate(msgs[:5])]))
    got = near.forward(ekd.SimpleFieldList(fields[:5]))
    for a, b in zip(got, ref):
        assert_same_values(a.to_numpy(), b.to_numpy(), "nearest on GRIB fields")


def test_grib_fields_of_the_wrong_grid_are_refused(cuda, tmp_path):
    from anemoi_transform_b200 import ekd
    from anemoi_transform_b200.filters import create_filter_by_name as F

    t_lat, t_lon = syn.octahedral(48)
    d, i, p, shape = syn.bilinear_matrix(2.0, t_lat, t_lon)
    path = str(tmp_path / "m.npz")
    syn.save_regrid_npz(path, d, i, p, shape, *syn.regular_latlon(2.0), t_lat, t_lon)
    fields = _fields(_messages(3, shape[1] - 5, [16], [0], [2]), shape[1] - 5)
    with pytest.raises(ValueError, match="dimension mismatch"):
        F("regrid", matrix=path).forward(ekd.SimpleFieldList(fields))


def test_float32_decode_feeds_the_fused_pipeline(cuda, tmp_path):
    """`set_decode_dtype(float32)`: values leave the unpack kernel as float32 (what
    `to_numpy(dtype=float32)` gives), a float32 matrix then returns float32 fields and the
    regrid | pointwise pipeline runs as ONE fused launch on GRIB input, with no host decode."""
    from anemoi_transform_b200 import ekd, grib
    from anemoi_transform_b200.filters import create_filter_by_name as F

    t_lat, t_lon = syn.octahedral(48)
    d, i, p, shape = syn.bilinear_matrix(2.0, t_lat, t_lon)
    s_lat, s_lon = syn.regular_latlon(2.0)
    path = str(tmp_path / "m.npz")
    syn.save_regrid_npz(path, d, i, p, shape, s_lat, s_lon, t_lat, t_lon)
    m = csr_array((d, i, p), shape=shape)
    n_src = shape[1]
    msgs = _messages(8, n_src, [16], [0], [2], seed=4)
    fields = _fields(msgs, n_src, s_lat, s_lon)
    x32 = [ogrib.decode(mm, n_points=n_src).astype(np.float32) for mm in msgs]
    grib.set_decode_dtype(np.float32)
    try:
        out = F("regrid", matrix=path).forward(ekd.SimpleFieldList(fields))
        for k, f in enumerate(out):
            got = f.to_numpy()
            assert got.dtype == np.float32
            assert_same_values(got, m @ x32[k], f"float32 field {k}")
        pipe = F("regrid", matrix=path) | F("clip_fields", param="t", minimum=250.0, maximum=290.0)
        out = pipe.forward(ekd.SimpleFieldList(fields))
        for k, f in enumerate(out):
            assert_same_values(f.to_numpy(), np.clip(m @ x32[k], np.float32(250.0), np.float32(290.0)), f"fused field {k}")
    finally:
        grib.set_decode_dtype(None)
    assert sum(f.decodes for f in fields) == 0


def test_pointwise_filters_take_grib_fields_without_a_host_decode(cuda):
    """uv_to_ddff and clip straight on GRIB-backed fields: values are decoded on the device
    (float64, like to_numpy()) and the result equals the same filter on the host-decoded fields."""
    from anemoi_transform_b200 import ekd
    from anemoi_transform_b200.filters import create_filter_by_name as F

    n = 40320
    rng = np.random.default_rng(21)
    lat, lon = syn.octahedral(96)
    msgs, md = [], []
    for lev in (500, 850):
        for param in ("u", "v"):
            msgs.append(ogrib.encode_grib2(rng.normal(0.0, 8.0, n), 16, 0))
            md.append(dict(param=param, levelist=lev, step=0))
    g_fields = [GribMessageField(m, n, d, latitudes=lat, longitudes=lon) for m, d in zip(msgs, md)]
    a_fields = [ekd.ArrayField(ogrib.decode(m), d, latitudes=lat, longitudes=lon) for m, d in zip(msgs, md)]
    for name, kw in (("uv_to_ddff", {}), ("clip_fields", dict(param="u", minimum=-3.0, maximum=4.0))):
        got = F(name, **kw).forward(ekd.SimpleFieldList(g_fields))
        assert sum(f.decodes for f in g_fields) == 0  # the filter itself decoded nothing on the host
        want = F(name, **kw).forward(ekd.SimpleFieldList(a_fields))
        assert [f.metadata("param") for f in got] == [f.metadata("param") for f in want]
        for a, b in zip(got, want):
            if a in g_fields:  # passed through untouched (clip of "u" leaves "v" alone): still the GRIB field
                continue
            assert a.to_numpy().dtype == b.to_numpy().dtype == np.float64
            assert_same_values(a.to_numpy(), b.to_numpy(), name)
    assert sum(f.decodes for f in g_fields) == 0


def test_mixed_fieldlist_splits_between_device_and_host_decode(cuda, tmp_path):
    """GRIB fields the device decodes (one of them with a bitmap: missing values -> NaN, which the
    matrix spreads to the targets that use them), a wrapper field and a plain numpy field in one
    FieldList: each takes its own route, the outputs keep the input order and equal scipy on the
    host-decoded values."""
    from anemoi_transform_b200 import ekd
    from anemoi_transform_b200.filters import create_filter_by_name as F

    t_lat, t_lon = syn.octahedral(48)
    d, i, p, shape = syn.bilinear_matrix(2.0, t_lat, t_lon)
    s_lat, s_lon = syn.regular_latlon(2.0)
    path = str(tmp_path / "m.npz")
    syn.save_regrid_npz(path, d, i, p, shape, s_lat, s_lon, t_lat, t_lon)
    m = csr_array((d, i, p), shape=shape)
    n_src = shape[1]
    rng = np.random.default_rng(31)
    msgs = _messages(6, n_src, [16, 12], [0], [2, 1], seed=31)
    fields = _fields(msgs, n_src, s_lat, s_lon)
    bm = rng.uniform(size=n_src) > 0.2
    with_bitmap = GribMessageField(ogrib.encode_grib2(rng.normal(285.0, 5.0, int(bm.sum())), 16, 0, bm), n_src, dict(param="sst", step=0), latitudes=s_lat, longitudes=s_lon)
    plain = ekd.ArrayField(rng.normal(0.0, 1.0, n_src).astype(np.float32), dict(param="z", step=0), latitudes=s_lat, longitudes=s_lon)
    from anemoi_transform_b200.fields import new_field_from_numpy

    doubled = rng.normal(1.0, 0.1, n_src)
    wrapper = new_field_from_numpy(doubled, template=fields[0], param="t2")  # forwards message() of the field it wraps
    mixed = fields[:3] + [with_bitmap, wrapper] + fields[3:] + [plain]
    out = F("regrid", matrix=path).forward(ekd.SimpleFieldList(mixed))
    assert [f.metadata("param") for f in out] == ["t"] * 3 + ["sst", "t2"] + ["t"] * 3 + ["z"]
    for f_in, f_out in zip(mixed, out):
        host = ogrib.decode(f_in.message(), n_points=n_src) if isinstance(f_in, GribMessageField) else np.asarray(f_in.to_numpy()).reshape(-1)
        assert_same_values(f_out.to_numpy(), m @ host, str(f_in.metadata("param")))
    assert np.isnan(out[3].to_numpy()).any() and not np.isnan(out[3].to_numpy()).all()
    assert sum(f.decodes for f in fields) == 0 and with_bitmap.decodes == 0
