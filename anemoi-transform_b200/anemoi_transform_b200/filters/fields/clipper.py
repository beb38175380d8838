"""`clip_fields` — reference `filters/fields/clipper.py:18-70`.

`np.clip(data, minimum, maximum)` on the selected param; NaN passes through; at least one
bound is required (`ValueError`, clipper.py:62).  Kernel: AT_EPI_PLAIN with AT_COL_CLIP_*.
"""

from __future__ import annotations

from typing import Any

from ... import _cabi
from ...filter import SingleFieldFilter
from . import filter_registry
from .pointwise import device_field, run_epilogue


@filter_registry.register("clip_fields")
class Clipper(SingleFieldFilter):
    """Clip the values of a single field to a specified range [minimum, maximum]."""

    required_inputs = ("param",)
    optional_inputs = {"minimum": None, "maximum": None}

    def prepare_filter(self) -> None:
        if self.minimum is None and self.maximum is None:
            raise ValueError("At least one value for minimum or maximum must be specified.")

    def forward_select(self) -> dict[str, Any]:
        return {"param": self.param}

    def _column(self) -> tuple:
        flags = (_cabi.COL_CLIP_LO if self.minimum is not None else 0) | (_cabi.COL_CLIP_HI if self.maximum is not None else 0)
        return (self.minimum or 0.0, self.maximum or 0.0, 0.0, flags)

    def forward_transform(self, field: Any) -> Any:
        return self.forward_transform_batch([field])[0]

    def forward_transform_batch(self, fields: list[Any]) -> list[Any]:
        out = run_epilogue(_cabi.EPI_PLAIN, fields, [self._column()] * len(fields))
        return [device_field(out, i, f, param=f.metadata("param")) for i, f in enumerate(fields)]
