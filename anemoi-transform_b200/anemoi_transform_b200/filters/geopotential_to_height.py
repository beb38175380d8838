"""`geopotential_to_height` (alias `orog_to_z`) and `height_to_geopotential` (alias `z_to_orog`)
— reference `filters/geopotential_to_height.py:19-56`.  Only the field branch is provided; the
tabular branch stays with the reference implementation.

The two spellings of the height key (`height`, `orography`) are mutually exclusive; the
geopotential key defaults to "z" and the height key to "orog".
"""

from __future__ import annotations

from typing import Any

from ..filter import DispatchingFilter
from . import filter_registry
from .fields.orog_to_z import Orography


def _height_key(config: dict[str, Any]) -> str:
    if "height" in config and "orography" in config:
        raise ValueError("Must specify either 'height' or 'orography' parameter, but not both.")
    return config.get("height", config.get("orography", "orog"))


class GeopotentialToHeight(DispatchingFilter):
    """Orography / height (m) ↔ geopotential (m²/s²) for field datasets."""

    def __init__(self, **config: Any) -> None:
        self.field_filter = Orography(geopotential=config.get("geopotential", "z"), orography=_height_key(config))

    def forward_fields(self, data: Any) -> Any:
        return self.field_filter.forward(data)

    def backward_fields(self, data: Any) -> Any:
        return self.field_filter.backward(data)


filter_registry.register("geopotential_to_height", GeopotentialToHeight, aliases=["orog_to_z"])
filter_registry.register("height_to_geopotential", GeopotentialToHeight.reversed, aliases=["z_to_orog"])
