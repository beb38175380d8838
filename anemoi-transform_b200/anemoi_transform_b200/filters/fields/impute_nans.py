"""`impute_nans_fields` — reference `filters/fields/impute_nans.py:22-55`.

NaNs of the selected fields become a fixed value (AT_EPI_IMPUTE_NAN); like the reference the
output is flattened (`to_numpy(flatten=True).copy()`, impute_nans.py:52).
"""

from __future__ import annotations

from typing import Any

from ... import _cabi
from ...fields import new_field_from_device_column
from ...filter import SingleFieldFilter
from . import filter_registry
from .pointwise import NO_COL, run_epilogue


@filter_registry.register("impute_nans_fields")
class ImputeNaNs(SingleFieldFilter):
    """A filter to impute NaN values in specified fields with a fixed value."""

    required_inputs = ("param", "value")

    def forward_select(self):
        return {"param": self.param}

    def forward_transform(self, field: Any) -> Any:
        return self.forward_transform_batch([field])[0]

    def forward_transform_batch(self, fields: list[Any]) -> list[Any]:
        out = run_epilogue(_cabi.EPI_IMPUTE_NAN, fields, [NO_COL] * len(fields), pa=float(self.value))
        return [new_field_from_device_column(*out.locate(i), template=f, shape=None) for i, f in enumerate(fields)]
