"""`geopotential_to_height` / `orog_to_z` dispatcher — reference `filters/geopotential_to_height.py:19-56`
(field branch; the tabular branch stays with the reference implementation)."""

from __future__ import annotations

from typing import Any

from ..filter import DispatchingFilter
from . import filter_registry
from .fields.orog_to_z import Orography as OrographyFields


class GeopotentialToHeight(DispatchingFilter):
    """Convert from geopotential to height for field datasets."""

    def __init__(self, **config: Any) -> None:
        config["geopotential"] = config.get("geopotential", "z")
        if ("height" in config) and ("orography" in config):
            raise ValueError("Must specify either 'height' or 'orography' parameter, but not both.")
        if "height" not in config:
            config["height"] = config.pop("orography", "orog")
        self.field_filter = OrographyFields(geopotential=config["geopotential"], orography=config["height"])

    def forward_fields(self, data: Any) -> Any:
        return self.field_filter.forward(data)

    def backward_fields(self, data: Any) -> Any:
        return self.field_filter.backward(data)


filter_registry.register("geopotential_to_height", GeopotentialToHeight, aliases=["orog_to_z"])
filter_registry.register("height_to_geopotential", GeopotentialToHeight.reversed, aliases=["z_to_orog"])
