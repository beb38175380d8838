"""Oracle: GRIB simple packing (grid-point data) — encoder and decoder in numpy.

TEST INFRASTRUCTURE — see oracle/__init__.py.

Where this sits on the path: the reference never touches GRIB bytes itself.  A GRIB-backed
FieldList reaches `RegridFilter.forward` as earthkit-data `GribField`s and the per-field
`field.to_numpy(flatten=True)` of `filters/fields/regrid.py:309` (and `matching.py:242-246`)
is where ecCodes decodes the message to float64 on one CPU core.  SURVEY §8(f) rank 4 lists
"optional GRIB-decode overlap" as the next data format on the input side of the hot path.

The arithmetic lives in un-vendored third parties (earthkit-data `>=0.12.4`, ecCodes behind it,
`pyproject.toml:39`; neither is installed here, no GRIB file ships with the reference's tests),
so this module restates the published algorithm:

* WMO Manual on Codes I.2, FM 92 GRIB edition 2: sections 0-8, data representation template
  5.0 ("grid point data - simple packing"), regulation 92.9.4:
      Y · 10^D = R + X · 2^E
  R IEEE float32, E and D sign-and-magnitude 16-bit integers, X unsigned big-endian bit field
  of `bitsPerValue` bits, values packed back to back, padded to an octet boundary.
* FM 92 GRIB edition 1: sections 0-5, BDS flag 0 (grid point, simple packing, float values),
  R in IBM System/360 hexadecimal float32, D in octets 27-28 of the PDS, plus ECMWF's
  long-message convention (3-octet lengths with the top bit set count units of 120 octets).
* ecCodes' order of operations in `data_simple_packing` (recalled from its sources, ecCodes
  2.3x `grib_accessor_class_data_simple_packing`): s = 2^E, d = 10^-D formed by |D| repeated
  multiplications / divisions by 10 in double precision starting from 1, every value
  `((X * s) + R) * d` in double precision, no fused multiply-add; bitsPerValue = 0 means a
  constant field equal to R.

PARITY UNPINNED: there is no golden vector for this step under /root/reference and ecCodes
cannot be run here.  What is checked instead (tests/test_oracle_grib.py): the encoder and the
decoder are written independently from the regulation above and invert each other within half
a packing unit for every bit width; known-answer bit patterns written out by hand (a
3-value 12-bit message, IBM and IEEE reference values) decode to their defining numbers.  For
D = 0 (what ECMWF's atmospheric fields use) the result does not depend on any recalled detail:
X · 2^E is exact and the one addition is correctly rounded.
"""

from __future__ import annotations

import math
import struct

import numpy as np


# ------------------------------------------------------------------------ scalars ------
def power(s: int, n: int) -> float:
    """n**s by repeated multiplication / division in double precision (ecCodes `codes_power`)."""
    v = 1.0
    if s == 0:
        return 1.0
    if s == 1:
        return float(n)
    while s < 0:
        v /= n
        s += 1
    while s > 0:
        v *= n
        s -= 1
    return v


def sign_magnitude16(v: int) -> bytes:
    return struct.pack(">H", (abs(v) & 0x7FFF) | (0x8000 if v < 0 else 0))


def from_sign_magnitude16(b: bytes) -> int:
    (u,) = struct.unpack(">H", b)
    return -(u & 0x7FFF) if u & 0x8000 else u & 0x7FFF


def ibm32_to_float(b: bytes) -> float:
    """IBM System/360 single: sign, 7-bit excess-64 base-16 exponent, 24-bit fraction."""
    (u,) = struct.unpack(">I", b)
    sign = -1.0 if u >> 31 else 1.0
    exponent = (u >> 24) & 0x7F
    mantissa = u & 0xFFFFFF
    if mantissa == 0:
        return 0.0 * sign
    return sign * mantissa * 16.0 ** (exponent - 64 - 6)


def float_to_ibm32_below(x: float) -> bytes:
    """The largest IBM single that is <= x (a reference value must not exceed the minimum)."""
    if x == 0.0:
        return b"\x00\x00\x00\x00"
    sign = 0x80 if x < 0 else 0
    a = abs(x)
    e = int(math.floor(math.log(a, 16))) + 1
    while a / 16.0**e >= 1.0:
        e += 1
    while a / 16.0**e < 1.0 / 16.0:
        e -= 1
    m = a / 16.0**e * 2**24
    m = math.floor(m) if x > 0 else math.ceil(m)  # round towards -inf
    if m >= 2**24:
        m //= 16
        e += 1
    return bytes([sign | ((e + 64) & 0x7F)]) + int(m).to_bytes(3, "big")


def float32_below(x: float) -> np.float32:
    r = np.float32(x)
    if float(r) > x:
        r = np.nextafter(r, np.float32(-np.inf))
    return r


# ------------------------------------------------------------------------ bit fields ---
def pack_bits(x: np.ndarray, nbits: int) -> bytes:
    """Unsigned integers → big-endian bit fields of `nbits` bits, back to back, zero padded."""
    if nbits == 0 or x.size == 0:
        return b""
    if nbits in (8, 16, 32):  # whole octets: a big-endian integer array is the bit stream
        return x.astype({8: ">u1", 16: ">u2", 32: ">u4"}[nbits]).tobytes()
    x = x.astype(np.uint64)
    shifts = np.arange(nbits - 1, -1, -1, dtype=np.uint64)
    bits = ((x[:, None] >> shifts[None, :]) & np.uint64(1)).astype(np.uint8).reshape(-1)
    return np.packbits(bits).tobytes()


def unpack_bits(buf: bytes, n: int, nbits: int) -> np.ndarray:
    if nbits == 0:
        return np.zeros(n, dtype=np.uint64)
    if nbits in (8, 16, 32):
        return np.frombuffer(buf, dtype={8: ">u1", 16: ">u2", 32: ">u4"}[nbits], count=n).astype(np.uint64)
    bits = np.unpackbits(np.frombuffer(buf, dtype=np.uint8), count=n * nbits).reshape(n, nbits).astype(np.uint64)
    weights = np.uint64(1) << np.arange(nbits - 1, -1, -1, dtype=np.uint64)
    return (bits * weights[None, :]).sum(axis=1, dtype=np.uint64)


# ------------------------------------------------------------------------ packing ------
def simple_packing_parameters(values: np.ndarray, nbits: int, decimal_scale: int, edition: int):
    """→ (R as Python float exactly representable in the edition's format, its 4 octets, E)."""
    scaled = values.astype(np.float64) * power(decimal_scale, 10)
    lo, hi = float(scaled.min()), float(scaled.max())
    if edition == 2:
        r32 = float32_below(lo)
        r, r_bytes = float(r32), struct.pack(">f", r32)
    else:
        r_bytes = float_to_ibm32_below(lo)
        r = ibm32_to_float(r_bytes)
    if nbits == 0 or hi == lo:
        return r, r_bytes, 0
    span = hi - r
    e = math.ceil(math.log2(span / (2.0**nbits - 1.0))) if span > 0 else 0
    while span / 2.0**e > 2.0**nbits - 1.0:
        e += 1
    return r, r_bytes, int(e)


def quantise(values: np.ndarray, r: float, e: int, nbits: int, decimal_scale: int) -> np.ndarray:
    if nbits == 0:
        return np.zeros(values.shape, dtype=np.uint64)
    scaled = values.astype(np.float64) * power(decimal_scale, 10)
    x = np.rint((scaled - r) / 2.0**e)
    return np.clip(x, 0, 2.0**nbits - 1).astype(np.uint64)


def encode_grib2(values: np.ndarray, nbits: int = 16, decimal_scale: int = 0, bitmap: np.ndarray | None = None) -> bytes:
    """One GRIB edition 2 message, template 5.0.  `bitmap` (bool per grid point) marks the points
    that carry a value; `values` then holds only those."""
    values = np.asarray(values, dtype=np.float64).reshape(-1)
    n_values = values.size
    n_points = n_values if bitmap is None else int(np.asarray(bitmap).size)
    const = n_values == 0 or float(values.max()) == float(values.min())
    nb = 0 if const else nbits
    r, r_bytes, e = simple_packing_parameters(values, nb, decimal_scale, 2) if n_values else (0.0, b"\0\0\0\0", 0)
    x = quantise(values, r, e, nb, decimal_scale)
    s1 = struct.pack(">IBHHBBBHBBBBBBB", 21, 1, 98, 0, 2, 0, 1, 2024, 1, 1, 0, 0, 0, 0, 1)
    # section 3: a template-less stub (grid definition template 65535 = "missing") of 14 octets + 58 reserved
    s3_body = struct.pack(">BIBBH", 0, n_points, 0, 0, 65535) + bytes(58)
    s3 = struct.pack(">IB", 5 + len(s3_body), 3) + s3_body
    s4_body = struct.pack(">HH", 0, 0) + bytes(25)
    s4 = struct.pack(">IB", 5 + len(s4_body), 4) + s4_body
    s5_body = struct.pack(">IH", n_values, 0) + r_bytes + sign_magnitude16(e) + sign_magnitude16(decimal_scale) + bytes([nb, 0])
    s5 = struct.pack(">IB", 5 + len(s5_body), 5) + s5_body
    if bitmap is None:
        s6 = struct.pack(">IBB", 6, 6, 255)
    else:
        bm = np.packbits(np.asarray(bitmap, dtype=bool).astype(np.uint8)).tobytes()
        s6 = struct.pack(">IBB", 6 + len(bm), 6, 0) + bm
    data = pack_bits(x, nb)
    s7 = struct.pack(">IB", 5 + len(data), 7) + data
    body = s1 + s3 + s4 + s5 + s6 + s7 + b"7777"
    s0 = b"GRIB" + bytes(2) + bytes([0, 2]) + struct.pack(">Q", 16 + len(body))
    return s0 + body


def encode_grib1(values: np.ndarray, nbits: int = 16, decimal_scale: int = 0, bitmap: np.ndarray | None = None) -> bytes:
    """One GRIB edition 1 message, BDS flag 0 (grid point, simple packing, float)."""
    values = np.asarray(values, dtype=np.float64).reshape(-1)
    n_values = values.size
    const = n_values == 0 or float(values.max()) == float(values.min())
    nb = 0 if const else nbits
    r, r_bytes, e = simple_packing_parameters(values, nb, decimal_scale, 1) if n_values else (0.0, b"\0\0\0\0", 0)
    x = quantise(values, r, e, nb, decimal_scale)
    pds = bytearray(28)
    pds[0:3] = (28).to_bytes(3, "big")
    pds[3], pds[4], pds[5], pds[6] = 128, 98, 255, 255
    pds[7] = 0x80 | (0x40 if bitmap is not None else 0)  # GDS present, BMS present?
    pds[8], pds[9] = 130, 100
    pds[26:28] = sign_magnitude16(decimal_scale)
    gds = bytearray(32)
    gds[0:3] = (32).to_bytes(3, "big")
    gds[4], gds[5] = 255, 255  # PV/PL absent, data representation type "missing"
    bms = b""
    if bitmap is not None:
        bits = np.packbits(np.asarray(bitmap, dtype=bool).astype(np.uint8)).tobytes()
        unused = len(bits) * 8 - int(np.asarray(bitmap).size)
        if (6 + len(bits)) % 2:
            bits += b"\0"
            unused += 8
        bms = (6 + len(bits)).to_bytes(3, "big") + bytes([unused]) + b"\0\0" + bits
    data = pack_bits(x, nb)
    unused4 = len(data) * 8 - n_values * nb
    if (11 + len(data)) % 2:  # sections have an even number of octets
        data += b"\0"
        unused4 += 8
    bds_len = 11 + len(data)
    bds_head = bytes([unused4 & 0x0F]) + sign_magnitude16(e) + r_bytes + bytes([nb])
    total = 8 + len(pds) + len(gds) + len(bms) + bds_len + 4
    if total < 0x800000:
        s0 = b"GRIB" + total.to_bytes(3, "big") + b"\x01"
        bds = bds_len.to_bytes(3, "big") + bds_head + data
    else:
        # ECMWF long messages: the 3-octet total counts units of 120 octets with the top bit set,
        # the stored section-4 length holds what must be subtracted: total = units * 120 - stored4;
        # the true section-4 length then follows from the total
        stored4 = (-total) % 120
        units = (total + stored4) // 120
        s0 = b"GRIB" + (0x800000 | units).to_bytes(3, "big") + b"\x01"
        bds = stored4.to_bytes(3, "big") + bds_head + data
    return s0 + bytes(pds) + bytes(gds) + bms + bds + b"7777"


# ------------------------------------------------------------------------ decoding -----
def scan(msg: bytes) -> dict:
    """Locate the packing parameters and the packed values of one message."""
    if msg[:4] != b"GRIB":
        raise ValueError("not a GRIB message")
    edition = msg[7]
    if edition == 2:
        (total,) = struct.unpack(">Q", msg[8:16])
        pos, info = 16, dict(edition=2, bitmap_offset=-1, has_bitmap=0, message_length=total)
        seen7 = 0
        while pos < total - 4:
            (length,) = struct.unpack(">I", msg[pos : pos + 4])
            number = msg[pos + 4]
            if number == 3:
                (info["n_points"],) = struct.unpack(">I", msg[pos + 6 : pos + 10])
            elif number == 5:
                (info["n_values"],) = struct.unpack(">I", msg[pos + 5 : pos + 9])
                (template,) = struct.unpack(">H", msg[pos + 9 : pos + 11])
                if template != 0:
                    raise NotImplementedError(f"data representation template 5.{template}")
                (info["reference_value"],) = struct.unpack(">f", msg[pos + 11 : pos + 15])
                info["reference_value"] = float(info["reference_value"])
                info["binary_scale"] = from_sign_magnitude16(msg[pos + 15 : pos + 17])
                info["decimal_scale"] = from_sign_magnitude16(msg[pos + 17 : pos + 19])
                info["bits_per_value"] = msg[pos + 19]
            elif number == 6:
                indicator = msg[pos + 5]
                if indicator == 0:
                    info["has_bitmap"], info["bitmap_offset"] = 1, pos + 6
                elif indicator != 255:
                    raise NotImplementedError(f"bitmap indicator {indicator}")
            elif number == 7:
                seen7 += 1
                info["data_offset"], info["data_length"] = pos + 5, length - 5
            pos += length
        if seen7 != 1:
            raise NotImplementedError("messages with several fields")
        if msg[total - 4 : total] != b"7777":
            raise ValueError("end section missing")
        return info
    if edition != 1:
        raise NotImplementedError(f"GRIB edition {edition}")
    len3 = int.from_bytes(msg[4:7], "big")
    pos = 8
    pds_len = int.from_bytes(msg[pos : pos + 3], "big")
    flag = msg[pos + 7]
    decimal_scale = from_sign_magnitude16(msg[pos + 26 : pos + 28])
    pos += pds_len
    if flag & 0x80:
        pos += int.from_bytes(msg[pos : pos + 3], "big")
    info = dict(edition=1, decimal_scale=decimal_scale, has_bitmap=0, bitmap_offset=-1)
    n_points = None
    if flag & 0x40:
        bms_len = int.from_bytes(msg[pos : pos + 3], "big")
        unused = msg[pos + 3]
        if int.from_bytes(msg[pos + 4 : pos + 6], "big") != 0:
            raise NotImplementedError("predefined bitmaps")
        info["has_bitmap"], info["bitmap_offset"] = 1, pos + 6
        n_points = (bms_len - 6) * 8 - unused
        pos += bms_len
    stored4 = int.from_bytes(msg[pos : pos + 3], "big")
    if len3 & 0x800000:
        total = (len3 & 0x7FFFFF) * 120 - stored4
        bds_len = total - 4 - pos
    else:
        total, bds_len = len3, stored4
    flags4 = msg[pos + 3]
    if flags4 & 0xF0:
        raise NotImplementedError("BDS flags: only grid-point simple packing of float values")
    unused4 = flags4 & 0x0F
    info["binary_scale"] = from_sign_magnitude16(msg[pos + 4 : pos + 6])
    info["reference_value"] = ibm32_to_float(msg[pos + 6 : pos + 10])
    nb = info["bits_per_value"] = msg[pos + 10]
    info["data_offset"], info["data_length"] = pos + 11, bds_len - 11
    info["n_values"] = ((bds_len - 11) * 8 - unused4) // nb if nb else -1
    if len3 & 0x800000 and nb:
        # the padding of a long message is not recorded in the unused-bits nibble
        info["n_values"] = -1
    info["n_points"] = n_points if n_points is not None else info["n_values"]
    info["message_length"] = total
    if msg[total - 4 : total] != b"7777":
        raise ValueError("end section missing")
    return info


def decode(msg: bytes, n_points: int | None = None, missing=np.nan) -> np.ndarray:
    """float64 values of one message, as `GribField.to_numpy(flatten=True)` gives them.

    `n_points` is needed where the message itself does not say (edition 1 constant fields and
    long messages).  Points masked out by a bitmap come back as `missing`."""
    info = scan(msg)
    n_values = info["n_values"]
    if info["has_bitmap"]:
        npts = info["n_points"] if info["n_points"] is not None and info["n_points"] >= 0 else n_points
        bits = np.unpackbits(np.frombuffer(msg, dtype=np.uint8, offset=info["bitmap_offset"]), count=npts).astype(bool)
        if n_values < 0:
            n_values = int(bits.sum())
    else:
        if n_values < 0:
            n_values = n_points
        bits = None
    nb = info["bits_per_value"]
    r = info["reference_value"]
    if nb == 0:
        vals = np.full(n_values, r, dtype=np.float64)
    else:
        s = power(info["binary_scale"], 2)
        d = power(-info["decimal_scale"], 10)
        if nb in (8, 16, 32):  # whole octets: one pass converts the big-endian integers (exactly)
            vals = np.frombuffer(msg, dtype={8: ">u1", 16: ">u2", 32: ">u4"}[nb], count=n_values, offset=info["data_offset"]).astype(np.float64)
        else:
            vals = unpack_bits(msg[info["data_offset"] : info["data_offset"] + info["data_length"]], n_values, nb).astype(np.float64)
        vals *= s  # ((X * s) + R) * d, one rounding each, as separate passes
        vals += r
        if d != 1.0:  # multiplying by 1.0 changes nothing
            vals *= d
    if bits is None:
        return vals
    out = np.full(bits.size, missing, dtype=np.float64)
    out[bits] = vals
    return out


# ------------------------------------------------------------------------ C port -------
def c_decode_regrid_f64(indptr, indices, data, messages: list, n_points: int, n_threads: int = 0) -> np.ndarray:
    """[F, n_tgt] float64: per message the 16-bit simple-packing decode and `matrix @ values`,
    fields spread over OpenMP threads (oracle/csr_matvec.c: `oracle_grib16_regrid_f64`) — the
    multi-core CPU baseline of bench.py's GRIB leg.  Bitwise equal to `matrix @ decode(msg)`."""
    import ctypes

    from . import spmm as ospmm

    lib = ospmm._c()
    fn = lib.oracle_grib16_regrid_f64
    fn.restype = None
    fn.argtypes = [ctypes.c_int64] + [ctypes.c_void_p] * 8 + [ctypes.c_int64] * 3 + [ctypes.c_int]
    indptr = np.ascontiguousarray(indptr, dtype=np.int32)
    indices = np.ascontiguousarray(indices, dtype=np.int32)
    data = np.ascontiguousarray(data, dtype=np.float32)
    infos = [scan(m) for m in messages]
    for i in infos:
        if i["bits_per_value"] != 16 or i["has_bitmap"]:
            raise NotImplementedError("the C baseline decodes 16-bit simple packing without a bitmap")
    views = [np.frombuffer(m, dtype=np.uint8) for m in messages]
    ptrs = (ctypes.c_void_p * len(messages))(*[v.ctypes.data + i["data_offset"] for v, i in zip(views, infos)])
    r = np.array([i["reference_value"] for i in infos], dtype=np.float64)
    s = np.array([power(i["binary_scale"], 2) for i in infos], dtype=np.float64)
    d = np.array([power(-i["decimal_scale"], 10) for i in infos], dtype=np.float64)
    n_tgt = indptr.shape[0] - 1
    out = np.empty((len(messages), n_tgt), dtype=np.float64)
    fn(n_tgt, indptr.ctypes.data, indices.ctypes.data, data.ctypes.data, ptrs, r.ctypes.data, s.ctypes.data, d.ctypes.data, out.ctypes.data, len(messages), n_points, n_tgt, int(n_threads))
    return out
