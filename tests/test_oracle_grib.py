"""GRIB simple packing: the numpy oracle against hand-written known answers and against itself
(encoder and decoder are independent restatements of WMO FM 92), and the C parser
(`at_grib_scan`, host only) against the oracle's."""

import ctypes
import struct

import numpy as np
import pytest

from oracle import grib as ogrib


def _grib2_by_hand(n_points, r_bytes, e, d, nbits, data, template=0):
    s1 = bytes.fromhex("00000015" "01" "0062" "0000" "02" "00" "01" "07e8" "01" "01" "00" "00" "00" "00" "01")
    s3 = struct.pack(">IBBIBBH", 72, 3, 0, n_points, 0, 0, 0) + bytes(58)
    s4 = struct.pack(">IB", 34, 4) + bytes(29)
    s5 = struct.pack(">IBIH", 21, 5, n_points, template) + r_bytes + struct.pack(">HHBB", e, d, nbits, 0)
    s6 = struct.pack(">IBB", 6, 6, 255)
    s7 = struct.pack(">IB", 5 + len(data), 7) + data
    body = s1 + s3 + s4 + s5 + s6 + s7 + b"7777"
    return b"GRIB" + bytes(2) + bytes([0, 2]) + struct.pack(">Q", 16 + len(body)) + body


def test_known_answer_edition2_12_bits():
    # R = 1.5 (0x3FC00000), E = -1 (sign-and-magnitude 0x8001), D = 0, X = 0x001, 0xABC, 0xFFF
    msg = _grib2_by_hand(3, bytes.fromhex("3fc00000"), 0x8001, 0, 12, bytes.fromhex("001abcfff0"))
    info = ogrib.scan(msg)
    assert (info["bits_per_value"], info["binary_scale"], info["decimal_scale"], info["reference_value"]) == (12, -1, 0, 1.5)
    assert info["n_points"] == 3 and info["n_values"] == 3 and info["data_length"] == 5
    np.testing.assert_array_equal(ogrib.decode(msg), [2.0, 1375.5, 2049.0])


def test_known_answer_decimal_scale_and_constant_field():
    # D = 2: values are (R + X) / 100 with 1/100 formed as 1/10/10; R = 250 (0x437A0000), 8 bits
    msg = _grib2_by_hand(2, bytes.fromhex("437a0000"), 0, 2, 8, bytes([0, 255]))
    d = 1.0 / 10.0 / 10.0
    np.testing.assert_array_equal(ogrib.decode(msg), [250.0 * d, 505.0 * d])
    const = _grib2_by_hand(4, bytes.fromhex("c0490fdb"), 0, 0, 0, b"")
    np.testing.assert_array_equal(ogrib.decode(const), np.full(4, float(np.float32(-3.1415927))))


def test_ibm_reference_values():
    assert ogrib.ibm32_to_float(bytes.fromhex("42640000")) == 100.0
    assert ogrib.ibm32_to_float(bytes.fromhex("c1100000")) == -1.0
    assert ogrib.ibm32_to_float(bytes.fromhex("00000000")) == 0.0
    assert ogrib.ibm32_to_float(bytes.fromhex("40800000")) == 0.5
    for x in (273.16, -12.75, 1e-5, 101325.0, 0.0):
        r = ogrib.ibm32_to_float(ogrib.float_to_ibm32_below(x))
        assert r <= x and (x == 0 or abs(r - x) <= abs(x) * 16 / 2**24)


@pytest.mark.parametrize("edition", [1, 2])
@pytest.mark.parametrize("nbits", [1, 3, 8, 11, 12, 16, 24, 32])
@pytest.mark.parametrize("decimal", [0, 2, -1])
def test_encode_decode_round_trip(edition, nbits, decimal):
    rng = np.random.default_rng(nbits * 10 + decimal + edition)
    v = rng.normal(280.0, 15.0, 1001)
    enc = ogrib.encode_grib2 if edition == 2 else ogrib.encode_grib1
    msg = enc(v, nbits, decimal)
    info = ogrib.scan(msg)
    assert info["bits_per_value"] == nbits and info["decimal_scale"] == decimal and info["edition"] == edition
    out = ogrib.decode(msg, n_points=v.size)
    unit = 2.0 ** info["binary_scale"] * 10.0 ** (-decimal)
    assert np.abs(out - v).max() <= 0.5000001 * unit + np.abs(v).max() * 1e-7
    # the packed integers survive a second trip exactly
    assert ogrib.decode(enc(out, nbits, decimal), n_points=v.size).shape == v.shape


def test_bitmap_long_message_and_constant():
    rng = np.random.default_rng(5)
    bm = rng.uniform(size=777) > 0.3
    v = rng.normal(0, 8, int(bm.sum()))
    for enc in (ogrib.encode_grib1, ogrib.encode_grib2):
        out = ogrib.decode(enc(v, 16, 0, bm))
        assert np.isnan(out[~bm]).all() and np.abs(out[bm] - v).max() < 1e-2
        assert (ogrib.decode(enc(np.full(100, 3.5), 16, 0), n_points=100) == 3.5).all()
    big = rng.normal(280, 15, 4_300_000)  # > 2^23 octets: ECMWF's long edition-1 message
    msg = ogrib.encode_grib1(big, 16, 0)
    info = ogrib.scan(msg)
    assert len(msg) > 0x800000 and info["message_length"] == len(msg) and info["n_values"] == -1
    assert np.abs(ogrib.decode(msg, n_points=big.size) - big).max() <= 2.0 ** info["binary_scale"] / 2 * 1.000001


# ------------------------------------------------------------------ the C parser ----------
def _c_scan(native_library, msg):
    from anemoi_transform_b200._cabi import GribInfo

    info = GribInfo()
    buf = np.frombuffer(msg, dtype=np.uint8)
    rc = native_library.at_grib_scan(ctypes.c_void_p(buf.ctypes.data), buf.size, ctypes.byref(info))
    return rc, info


@pytest.mark.parametrize("edition", [1, 2])
def test_c_parser_equals_the_oracle(native_library, edition):
    rng = np.random.default_rng(edition)
    enc = ogrib.encode_grib2 if edition == 2 else ogrib.encode_grib1
    cases = [(enc(rng.normal(280, 15, n), nb, d), n) for nb in (1, 7, 8, 12, 16, 24, 32) for d in (0, 3, -2) for n in (1, 9, 1001)]
    bm = rng.uniform(size=300) > 0.5
    cases.append((enc(rng.normal(0, 1, int(bm.sum())), 16, 0, bm), 300))
    cases.append((enc(np.full(64, -7.25), 16, 0), 64))
    if edition == 1:
        cases.append((ogrib.encode_grib1(rng.normal(280, 15, 4_300_000), 16, 0), 4_300_000))
    for msg, _n in cases:
        rc, c = _c_scan(native_library, msg)
        assert rc == 0, native_library.at_last_error()
        o = ogrib.scan(msg)
        for key in ("edition", "bits_per_value", "binary_scale", "decimal_scale", "has_bitmap", "reference_value", "data_offset", "data_length", "bitmap_offset", "message_length"):
            assert getattr(c, key) == o[key], (key, getattr(c, key), o[key])
        assert c.n_values == o["n_values"]
        assert c.n_points == (o["n_points"] if o["n_points"] is not None else -1)


def test_c_parser_refuses_what_the_device_does_not_decode(native_library):
    from anemoi_transform_b200 import _cabi, grib

    ok = _grib2_by_hand(3, bytes.fromhex("3fc00000"), 0x8001, 0, 12, bytes.fromhex("001abcfff0"))
    assert _c_scan(native_library, ok)[0] == 0
    ccsds = _grib2_by_hand(3, bytes.fromhex("3fc00000"), 0x8001, 0, 12, bytes.fromhex("001abcfff0"), template=42)
    assert _c_scan(native_library, ccsds)[0] == _cabi.AT_ERR_UNSUPPORTED and b"5.42" in native_library.at_last_error()
    assert _c_scan(native_library, ok[:-10])[0] == _cabi.AT_ERR_INVALID  # truncated
    assert _c_scan(native_library, b"NOPE" + ok[4:])[0] == _cabi.AT_ERR_INVALID
    assert _c_scan(native_library, ok[:7] + b"\x03" + ok[8:])[0] == _cabi.AT_ERR_UNSUPPORTED  # edition 3
    spectral = bytearray(ogrib.encode_grib1(np.arange(10.0), 16, 0))
    bds = ogrib.scan(bytes(spectral))["data_offset"] - 11
    spectral[bds + 3] |= 0x80  # spherical harmonics flag
    assert _c_scan(native_library, bytes(spectral))[0] == _cabi.AT_ERR_UNSUPPORTED
    # the Python front end turns all of those into "take the to_numpy() route"
    assert grib.scan(ccsds) is None and grib.scan(ok[:-10]) is None and grib.scan(ok) is not None


def test_only_fields_that_own_a_message_take_the_packed_route(native_library):
    from grib_fields import GribMessageField

    from anemoi_transform_b200 import grib
    from anemoi_transform_b200.fields import new_field_from_numpy

    rng = np.random.default_rng(0)
    v = rng.normal(280, 15, 500)
    f = GribMessageField(ogrib.encode_grib2(v, 16), 500, {"param": "t"})
    packed = grib.packed_of([f, f])
    assert packed is not None and packed.n_fields == 2 and packed.n_points == 500 and packed.packed_bytes == 2000
    # a wrapper forwards message() to the field it wraps, whose data it replaced: not packed
    wrapped = new_field_from_numpy(v * 2, template=f)
    assert wrapped.message() == f.message() and grib.packed_of([wrapped]) is None
    # a bitmap is fine (missing points decode to NaN on the device) as long as it covers the grid
    bm = rng.uniform(size=500) > 0.5
    with_bitmap = GribMessageField(ogrib.encode_grib2(v[bm], 16, 0, bm), 500, {"param": "sst"})
    both = grib.packed_of([f, with_bitmap])
    assert both is not None and both.n_points == 500 and both.packed_bytes == 1000 + 2 * int(bm.sum()) + 63
    short_bitmap = GribMessageField(ogrib.encode_grib2(v[:400][bm[:400]], 16, 0, bm[:400]), 500, {"param": "sst"})
    assert grib.packed_of([short_bitmap]) is None
    # mixed sizes, other packings and plain numpy fields keep the to_numpy() route; split() sorts them out
    other = GribMessageField(ogrib.encode_grib2(v[:400], 16), 400, {"param": "t"})
    assert grib.packed_of([f, other]) is None
    packed, packed_idx, rest = grib.split([f, wrapped, with_bitmap, other])
    assert packed is not None and packed_idx == [0, 2] and rest == [1, 3]
    import os

    os.environ["AT_B200_GRIB_DEVICE_DECODE"] = "0"
    try:
        assert grib.packed_of([f]) is None
    finally:
        del os.environ["AT_B200_GRIB_DEVICE_DECODE"]


def test_c_decode_regrid_equals_numpy_decode_then_scipy():
    from scipy.sparse import csr_array

    from anemoi_transform_b200 import synthetic as syn

    t_lat, t_lon = syn.octahedral(24)
    d, i, p, shape = syn.bilinear_matrix(2.0, t_lat, t_lon)
    m = csr_array((d, i, p), shape=shape)
    rng = np.random.default_rng(2)
    msgs = [ogrib.encode_grib2(rng.normal(280, 15, shape[1]), 16, dec) for dec in (0, 0, 2, -1, 0)]
    got = ogrib.c_decode_regrid_f64(p, i, d, msgs, shape[1], n_threads=3)
    for k, msg in enumerate(msgs):
        want = m @ ogrib.decode(msg)
        assert want.dtype == np.float64 and np.array_equal(got[k], want)
