"""`remove_nans_fields` — reference `filters/fields/remove_nans.py:24-119`.

Grid points where the first field (or the field named by `param`) is NaN are dropped from
every field and from the grid.  Device path: `at_compare_mask` ("is not NaN") →
`at_compact_mask` (sorted indices of the kept points) → one `at_gather_rows` over the whole
batch.  The mask is computed once and reused by later calls, like the reference (95-107).
"""

from __future__ import annotations

import logging
from typing import Any

from ... import _cabi
from ...batching import fields_to_batch
from ...device import DeviceBatch, compact_mask, compare_mask, gather_rows, results_are_host_bound
from ...fields import new_field_from_device_column, new_field_from_latitudes_longitudes, new_fieldlist_from_list
from ...filter import Filter
from . import filter_registry

LOG = logging.getLogger(__name__)


@filter_registry.register("remove_nans_fields")
class RemoveNaNs(Filter):
    """A filter to mask out NaNs."""

    def __init__(self, *, method: str = "mask", check: bool = False, param: str | None = None):
        self.method = method
        self.check = check
        self.param = param
        assert method == "mask", f"Method {method} not implemented"
        assert not check, "Check not implemented"
        self._mask = None  # host bool array, like the reference's attribute
        self._index = None  # device int64 indices of the kept points
        self._latitudes = None
        self._longitudes = None

    def forward(self, fields: Any) -> Any:
        fields = list(fields)
        if not fields:
            return new_fieldlist_from_list([])
        batch = fields_to_batch(fields)
        if self._mask is None:
            if self.param is None:
                first_i = 0
            else:
                for first_i, first in enumerate(fields):
                    if first.metadata("param") == self.param:
                        break
                else:
                    raise ValueError(f"{self.param=} not found in\n{[f.metadata('param') for f in fields]}")
            first = fields[first_i]
            keep = compare_mask(batch.data[:, first_i], _cabi.CMP_NOT_NAN, 0.0)
            self._index = compact_mask(keep)
            self._mask = keep.cpu().numpy().astype(bool)
            latitudes, longitudes = first.grid_points()
            self._latitudes = latitudes[self._mask]
            self._longitudes = longitudes[self._mask]
        if batch.n_points != self._mask.shape[0]:
            raise IndexError(f"boolean index did not match indexed array along axis 0; size of axis is {batch.n_points} but size of corresponding boolean axis is {self._mask.shape[0]}")
        out = DeviceBatch(gather_rows(batch.data, self._index, n_fields=batch.n_fields), batch.n_fields)
        if results_are_host_bound():
            out.prefetch()
        return new_fieldlist_from_list(
            [
                new_field_from_latitudes_longitudes(
                    new_field_from_device_column(out, i, template=f, shape=None), latitudes=self._latitudes, longitudes=self._longitudes
                )
                for i, f in enumerate(fields)
            ]
        )
