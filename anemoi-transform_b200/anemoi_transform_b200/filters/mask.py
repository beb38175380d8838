"""`mask` / `apply_mask` dispatcher — reference `filters/mask.py:19-35`."""

from __future__ import annotations

from typing import Any

from ..filter import DispatchingFilter
from . import filter_registry
from .fields.apply_mask import MaskVariable as MaskVariableFields


class Mask(DispatchingFilter):
    """Mask field datasets (the tabular branch of the reference is outside this package)."""

    def __init__(self, **config: Any) -> None:
        if "path" in config or "mask_param" in config:
            self.filter = MaskVariableFields(**config)
        else:
            raise NotImplementedError(
                "mask: only the field form (`path` or `mask_param`) is provided; "
                "tabular masking stays with the reference implementation"
            )

    def forward_fields(self, data: Any) -> Any:
        return self.filter.forward(data)


filter_registry.register("mask", Mask, aliases=["apply_mask"])
