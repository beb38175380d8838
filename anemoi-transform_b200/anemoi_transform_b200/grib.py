"""GRIB-backed fields: hand the packed message to the device instead of decoding on the host.

In the reference a GRIB FieldList reaches a filter as earthkit-data `GribField`s and every
`field.to_numpy(flatten=True)` (`filters/fields/regrid.py:309`, `matching.py:242-246`) has
ecCodes decode one message to float64 on one core.  A field whose *class* provides
`message()` (earthkit-data's `GribField.message()` returns the encoded message) and whose
message is grid-point simple packing without a bitmap is uploaded packed — 2 bytes per point
at 16 bits instead of 8 — and decoded by `grib_unpack_kernel` straight into the
[points x fields] batch (`csrc/grib.cu`).  Everything else keeps the `to_numpy()` route.

Wrapper fields (`NewDataField` and friends) forward unknown attributes to the field they
wrap, so `message()` of a wrapper would return the *original* data; only a `message` defined
on the field's own class counts.

`AT_B200_GRIB_DEVICE_DECODE=0` switches the packed route off.
"""

from __future__ import annotations

import os
from ctypes import byref, c_void_p
from typing import Any, Sequence

import numpy as np

from . import _cabi
from ._cabi import GribInfo


def enabled() -> bool:
    return os.environ.get("AT_B200_GRIB_DEVICE_DECODE", "1") != "0"


def scan(message) -> GribInfo | None:
    """Packing parameters of one message, or None when it is not something the device decodes
    (not GRIB, another packing, several fields).  Host only."""
    buf = np.frombuffer(message, dtype=np.uint8)
    info = GribInfo()
    try:
        _cabi.call("at_grib_scan", c_void_p(buf.ctypes.data), buf.size, byref(info))
    except (_cabi.NativeCallError, ValueError):
        return None
    return info


def _message_of(field: Any):
    method = getattr(type(field), "message", None)
    if method is None or not callable(method):
        return None
    try:
        m = field.message()
    except Exception:
        return None
    if isinstance(m, (bytes, bytearray, memoryview)) and len(m) >= 16:
        return m
    return None


def is_packed_message(field: Any) -> bool:
    """Would `packed_of([field])` take the packed route?"""
    return packed_of([field]) is not None


class PackedFields:
    """The messages of a FieldList the device can decode, with their scans."""

    def __init__(self, buffers: list[np.ndarray], infos, n_points: int):
        self.buffers = buffers  # uint8 views on the messages: keep them alive while in use
        self.infos = infos  # ctypes array of GribInfo
        self.n_points = n_points
        self.n_fields = len(buffers)
        self.pointers = (c_void_p * self.n_fields)(*[b.ctypes.data for b in buffers])

    #: what `GribField.to_numpy()` returns
    dtype = np.dtype(np.float64)

    @property
    def packed_bytes(self) -> int:
        return int(sum((self.n_points * i.bits_per_value + 7) // 8 for i in self.infos))

    def slice(self, lo: int, hi: int) -> "PackedFields":
        infos = (GribInfo * (hi - lo))(*self.infos[lo:hi])
        return PackedFields(self.buffers[lo:hi], infos, self.n_points)


def packed_of(fields: Sequence[Any]) -> PackedFields | None:
    """`PackedFields` when *every* field is a simple-packed GRIB message of the same grid size
    without a bitmap, else None (the caller then uses `to_numpy()` for all of them)."""
    if not fields or not enabled():
        return None
    buffers: list[np.ndarray] = []
    scans: list[GribInfo] = []
    n_points = -1
    for f in fields:
        m = _message_of(f)
        if m is None:
            return None
        info = scan(m)
        if info is None or info.has_bitmap:
            return None
        n = info.n_values if info.n_values >= 0 else info.n_points
        shape = getattr(f, "shape", None)
        if shape is not None:
            declared = int(np.prod(shape))
            if n >= 0 and n != declared:
                return None
            n = declared
        if n < 0 or (n_points >= 0 and n != n_points):
            return None
        if info.data_length < (n * info.bits_per_value + 7) // 8:
            return None
        n_points = n
        buffers.append(np.frombuffer(m, dtype=np.uint8))
        scans.append(info)
    return PackedFields(buffers, (GribInfo * len(scans))(*scans), n_points)


def upload(packed: PackedFields, dtype=np.float64):
    """→ `DeviceBatch` holding the decoded fields as columns (`at_hostio_upload_grib`)."""
    from .device import AT_F32, AT_F64, DeviceBatch, HostIO, _ptr, empty_batch, require_cuda, stream_ptr

    torch = require_cuda()
    dtype = np.dtype(dtype)
    tdtype = torch.float32 if dtype == np.float32 else torch.float64
    pm = empty_batch(packed.n_points, packed.n_fields, tdtype, "cuda")
    _cabi.call(
        "at_hostio_upload_grib",
        HostIO.get().handle,
        packed.pointers,
        packed.infos,
        packed.n_fields,
        packed.n_points,
        AT_F32 if dtype == np.float32 else AT_F64,
        _ptr(pm),
        int(pm.stride(0)),
        stream_ptr(),
    )
    return DeviceBatch(pm, packed.n_fields)


__all__ = ["GribInfo", "PackedFields", "enabled", "is_packed_message", "packed_of", "scan", "upload"]
