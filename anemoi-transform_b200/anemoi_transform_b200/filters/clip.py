"""`clip` / `clipper` dispatcher — reference `filters/clip.py:19-35`."""

from __future__ import annotations

from typing import Any

from ..filter import DispatchingFilter
from . import filter_registry
from .fields.clipper import Clipper as ClipperFields


class Clip(DispatchingFilter):
    """Clip field datasets (the tabular branch of the reference is outside this package)."""

    def __init__(self, **config: Any) -> None:
        if "param" in config and isinstance(config["param"], str):
            self.filter = ClipperFields(**config)
        else:
            raise NotImplementedError(
                "clip: only the field form (`param: <str>`, minimum / maximum) is provided; "
                "tabular clipping stays with the reference implementation"
            )

    def forward_fields(self, data: Any) -> Any:
        return self.filter.forward(data)


filter_registry.register("clip", Clip, aliases=["clipper"])
