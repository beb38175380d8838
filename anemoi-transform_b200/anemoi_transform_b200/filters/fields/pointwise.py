"""Shared plumbing of the pointwise field filters: build an epilogue program over a
point-major batch, run it through `at_pointwise`, hand back device-column fields."""

from __future__ import annotations

from typing import Any, Sequence

import numpy as np

from ... import _cabi
from ...batching import fields_to_batch, numpy_dtype_of
from ...device import DeviceBatch, Epilogue, results_are_host_bound, round_up
from ...fields import device_column_of, new_field_from_device_column

NO_COL = (0.0, 0.0, 0.0, 0)  # (lo, hi, pressure, flags) of an output column with nothing to do

OUT_PER_GROUP = _cabi.EPI_OUT_PER_GROUP


#: kinds whose inputs come in partner pairs (u with v, q with t, …); the others are one-to-one
PAIR_KINDS = frozenset(
    {
        _cabi.EPI_UV2DDFF, _cabi.EPI_DDFF2UV, _cabi.EPI_QT2R, _cabi.EPI_QT2QTR, _cabi.EPI_RT2Q, _cabi.EPI_RT2RTQ,
        _cabi.EPI_ATAN2, _cabi.EPI_RT2D, _cabi.EPI_RT2RTD, _cabi.EPI_DT2R, _cabi.EPI_DT2DTR,
    }
)  # fmt: skip


class SplitOutput:
    """Outputs of an epilogue that ran once per dtype run: output column j lives in
    `parts[k]` at a local column (`locate`)."""

    def __init__(self, where: list[tuple[DeviceBatch, int]]):
        self._where = where

    def locate(self, col: int) -> tuple[DeviceBatch, int]:
        return self._where[col]


def _float_dtype(dt) -> np.dtype:
    return np.dtype(np.float32) if dt == np.float32 else np.dtype(np.float64)


def run_epilogue(kind: int, inputs: Sequence[Any], out_cols: Sequence[tuple], row_mask=None, pa: float = 0.0, pb: float = 0.0, batch: DeviceBatch | None = None):
    """Run one uniform-kind epilogue over `inputs` (fields in partner order).

    `out_cols[j]` = (lo, hi, pressure, flags) of output column j (real columns only; padding
    is added here); `pa`, `pb` are the kind's constants.  Returns the output batch — `locate(j)`
    names the batch and column of output j.  `batch`, when given, is `fields_to_batch(inputs)`
    already built.

    numpy keeps each field's dtype (a float32 field stays float32 next to a float64 one), so
    a FieldList of mixed dtypes runs once per run of equal dtype rather than being promoted to
    the widest (partners of one group share the group's `np.result_type`)."""
    if batch is not None:
        return _run_uniform(kind, inputs, out_cols, row_mask, pa, pb, batch)
    cols = [device_column_of(f) for f in inputs]
    if inputs and all(c is None for c in cols):
        from ... import grib

        packed = grib.packed_of(inputs)
        if packed is not None:  # GRIB messages: decoded on the device, one dtype for all, no host decode
            return _run_uniform(kind, inputs, out_cols, row_mask, pa, pb, grib.upload(packed))
    values = [None if c is not None else np.asarray(f.to_numpy()).reshape(-1) for f, c in zip(inputs, cols)]
    dtypes = [_float_dtype(numpy_dtype_of(c[0]) if c is not None else v.dtype) for c, v in zip(cols, values)]
    g_in = 2 if kind in PAIR_KINDS else 1
    g_out = OUT_PER_GROUP[kind] * g_in // 4
    n_groups = len(inputs) // g_in
    group_dtype = [np.result_type(*dtypes[g * g_in : (g + 1) * g_in]) for g in range(n_groups)]
    if len(set(group_dtype)) <= 1:
        host = values if all(v is not None for v in values) else None
        return _run_uniform(kind, inputs, out_cols, row_mask, pa, pb, fields_to_batch(inputs, host_values=host))
    where: list[tuple[DeviceBatch, int]] = [None] * len(out_cols)  # type: ignore[list-item]
    for dt in dict.fromkeys(group_dtype):
        groups = [g for g in range(n_groups) if group_dtype[g] == dt]
        members = [i for g in groups for i in range(g * g_in, (g + 1) * g_in)]
        part_inputs = [inputs[i] for i in members]
        host = [values[i] if values[i] is not None and values[i].dtype == dt else None for i in members]
        if any(h is None for h, i in zip(host, members) if cols[i] is None):
            host = [np.ascontiguousarray(np.asarray(inputs[i].to_numpy()).reshape(-1), dtype=dt) for i in members]
        sub = fields_to_batch(part_inputs, host_values=host if all(h is not None for h in host) else None)
        outs = [j for g in groups for j in range(g * g_out, (g + 1) * g_out)]
        part = _run_uniform(kind, part_inputs, [out_cols[j] for j in outs if j < len(out_cols)], row_mask, pa, pb, sub)
        for k, j in enumerate(outs):
            if j < len(out_cols):
                where[j] = (part, k)
    return SplitOutput(where)


def _run_uniform(kind: int, inputs: Sequence[Any], out_cols: Sequence[tuple], row_mask, pa: float, pb: float, batch: DeviceBatch) -> DeviceBatch:
    n_in = round_up(len(inputs), 4)
    n_out_real = len(out_cols)
    n_out = n_in // 4 * OUT_PER_GROUP[kind]
    cols = list(out_cols) + [NO_COL] * (n_out - n_out_real)
    epi = Epilogue([(kind, 0, n_in, 0, pa, pb)], cols)
    try:
        out = epi.apply(batch.data, row_mask=row_mask)
    finally:
        epi.close()
    result = DeviceBatch(out, n_out_real)
    if results_are_host_bound():
        result.prefetch()
    return result


def device_field(batch: Any, col: int, template: Any, **metadata: Any) -> Any:
    batch, col = batch.locate(col)
    return new_field_from_device_column(batch, col, template=template, shape=getattr(template, "shape", None), **metadata)
