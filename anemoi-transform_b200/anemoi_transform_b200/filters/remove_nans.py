"""`remove_nans` / `drop_nans` dispatcher — reference `filters/remove_nans.py:19-50`."""

from __future__ import annotations

from typing import Any

from ..filter import DispatchingFilter
from . import filter_registry
from .fields.remove_nans import RemoveNaNs as RemoveNaNsFields


class RemoveNaNs(DispatchingFilter):
    """Remove NaNs in field datasets (the tabular branch of the reference is outside this package)."""

    def __init__(self, **config: Any) -> None:
        if ("columns" in config) or ("column_prefix" in config) or ("how" in config):
            raise NotImplementedError("remove_nans: the tabular form (`columns` / `column_prefix` / `how`) stays with the reference implementation")
        self.field_filter = RemoveNaNsFields(**config)

    def forward_fields(self, data: Any) -> Any:
        return self.field_filter.forward(data)


filter_registry.register("remove_nans", RemoveNaNs, aliases=["drop_nans"])
