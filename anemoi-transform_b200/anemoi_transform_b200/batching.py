"""Moving FieldList values in and out of point-major device batches."""

from __future__ import annotations

from typing import Any, Sequence

import numpy as np

from .device import DeviceBatch, empty_batch, gather_cols, require_cuda, round_up
from .fields import device_column_of


def fields_to_batch(fields: Sequence[Any], host_values: Sequence[Any] | None = None) -> DeviceBatch:
    """A point-major batch whose column j holds the values of fields[j].

    Fields already resident in HBM (outputs of an earlier filter of this package) are not
    round-tripped through the host: consecutive columns of one batch are used in place,
    other arrangements are re-packed on the device.  `host_values[j]`, when given, is
    fields[j].to_numpy(flatten=True) already fetched by the caller.
    """
    torch = require_cuda()
    cols = [device_column_of(f) for f in fields]
    if cols and all(c is not None for c in cols):
        batches = {id(c[0]) for c in cols}
        first_batch, c0 = cols[0]
        n = len(cols)
        consecutive = len(batches) == 1 and all(c[1] == c0 + j for j, c in enumerate(cols))
        align = 4 if first_batch.data.element_size() == 4 else 2
        if consecutive and c0 % align == 0 and c0 + round_up(n, 4) <= first_batch.data.shape[1]:
            return DeviceBatch(first_batch.data[:, c0 : c0 + round_up(n, 4)], n)
        dtypes = {c[0].data.dtype for c in cols}
        if len(dtypes) == 1 and len({c[0].n_points for c in cols}) == 1:
            out = empty_batch(first_batch.n_points, n, first_batch.data.dtype, first_batch.data.device)
            # one column-gather launch per source batch, each writing its own output columns
            by_batch: dict[int, list[tuple[int, int]]] = {}
            for j, (b, c) in enumerate(cols):
                by_batch.setdefault(id(b), []).append((j, c))
            lookup = {id(c[0]): c[0] for c in cols}
            if len(by_batch) == 1:
                gather_cols(first_batch.data, [c[1] for c in cols], out=out)
            else:
                for key, pairs in by_batch.items():
                    tmp = gather_cols(lookup[key].data, [c for _, c in pairs])
                    # scatter this batch's columns to their places (contiguous runs in practice)
                    for k, (j, _) in enumerate(pairs):
                        out[:, j] = tmp[:, k]
            return DeviceBatch(out, n)
    if host_values is not None and all(v is not None for v in host_values):
        return DeviceBatch.from_host_fields(list(host_values))
    if all(c is None for c in cols):
        from . import grib

        packed = grib.packed_of(fields)
        if packed is not None:  # GRIB messages: packed octets up, decoded on the device
            return grib.upload(packed)
    return DeviceBatch.from_host_fields([np.asarray(f.to_numpy()).reshape(-1) for f in fields])


def numpy_dtype_of(batch: DeviceBatch):
    torch = require_cuda()
    return np.float32 if batch.data.dtype == torch.float32 else np.float64
