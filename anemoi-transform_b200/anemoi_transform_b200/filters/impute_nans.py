"""`impute_nans` / `replace_nans` dispatcher — reference `filters/impute_nans.py:19-50`."""

from __future__ import annotations

from typing import Any

from ..filter import DispatchingFilter
from . import filter_registry
from .fields.impute_nans import ImputeNaNs as ImputeNaNsFields


class ImputeNaNs(DispatchingFilter):
    """Impute NaN values in field datasets (the tabular branch of the reference is outside this package)."""

    def __init__(self, **config: Any) -> None:
        if ("columns" in config) or ("column_prefix" in config):
            raise NotImplementedError("impute_nans: the tabular form (`columns` / `column_prefix`) stays with the reference implementation")
        self.field_filter = ImputeNaNsFields(**config)

    def forward_fields(self, data: Any) -> Any:
        return self.field_filter.forward(data)


filter_registry.register("impute_nans", ImputeNaNs, aliases=["replace_nans"])
