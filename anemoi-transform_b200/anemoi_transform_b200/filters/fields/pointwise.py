"""Shared plumbing of the pointwise field filters: build an epilogue program over a
point-major batch, run it through `at_pointwise`, hand back device-column fields."""

from __future__ import annotations

from typing import Any, Sequence

from ... import _cabi
from ...batching import fields_to_batch
from ...device import DeviceBatch, Epilogue, round_up
from ...fields import new_field_from_device_column

NO_COL = (0.0, 0.0, 0.0, 0)  # (lo, hi, pressure, flags) of an output column with nothing to do

OUT_PER_GROUP = _cabi.EPI_OUT_PER_GROUP


def run_epilogue(kind: int, inputs: Sequence[Any], out_cols: Sequence[tuple], row_mask=None, pa: float = 0.0, pb: float = 0.0, batch: DeviceBatch | None = None) -> DeviceBatch:
    """Run one uniform-kind epilogue over `inputs` (fields in partner order).

    `out_cols[j]` = (lo, hi, pressure, flags) of output column j (real columns only; padding
    is added here); `pa`, `pb` are the kind's constants.  Returns the output batch; its
    column j is output j.  `batch`, when given, is `fields_to_batch(inputs)` already built.
    """
    if batch is None:
        batch = fields_to_batch(inputs)
    n_in = round_up(len(inputs), 4)
    n_out_real = len(out_cols)
    n_out = n_in // 4 * OUT_PER_GROUP[kind]
    cols = list(out_cols) + [NO_COL] * (n_out - n_out_real)
    epi = Epilogue([(kind, 0, n_in, 0, pa, pb)], cols)
    try:
        out = epi.apply(batch.data, row_mask=row_mask)
    finally:
        epi.close()
    return DeviceBatch(out, n_out_real)


def device_field(batch: DeviceBatch, col: int, template: Any, **metadata: Any) -> Any:
    return new_field_from_device_column(batch, col, template=template, shape=getattr(template, "shape", None), **metadata)
