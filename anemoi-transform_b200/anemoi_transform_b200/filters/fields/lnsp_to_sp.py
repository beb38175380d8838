"""`lnsp_to_sp` / `sp_to_lnsp` — reference `filters/fields/lnsp_to_sp.py:19-103`.

exp / log of every selected field in one device pass (AT_EPI_EXP / AT_EPI_LOG).
"""

from __future__ import annotations

from typing import Any

from ... import _cabi
from ...filter import SingleFieldFilter
from . import filter_registry
from ._requests import swap_param
from .pointwise import NO_COL, device_field, run_epilogue


class LnspToSp(SingleFieldFilter):
    """A filter to convert natural log of surface pressure (lnsp) to surface pressure (sp), and back."""

    optional_inputs = {"log_of_surface_pressure": "lnsp", "surface_pressure": "sp"}

    def forward_select(self):
        return {"param": self.log_of_surface_pressure}

    def backward_select(self):
        return {"param": self.surface_pressure}

    def forward_transform(self, log_of_surface_pressure: Any) -> Any:
        return self.forward_transform_batch([log_of_surface_pressure])[0]

    def backward_transform(self, surface_pressure: Any) -> Any:
        return self.backward_transform_batch([surface_pressure])[0]

    def forward_transform_batch(self, fields: list[Any]) -> list[Any]:
        out = run_epilogue(_cabi.EPI_EXP, fields, [NO_COL] * len(fields))
        new_metadata = {"param": self.surface_pressure, "levelist": None, "level": None}
        return [device_field(out, i, f, **new_metadata) for i, f in enumerate(fields)]

    def backward_transform_batch(self, fields: list[Any]) -> list[Any]:
        out = run_epilogue(_cabi.EPI_LOG, fields, [NO_COL] * len(fields))
        return [device_field(out, i, f, param=self.log_of_surface_pressure) for i, f in enumerate(fields)]

    def patch_data_request(self, data_request: dict[str, Any]) -> dict[str, Any]:
        return swap_param(
            data_request,
            self.surface_pressure,
            self.log_of_surface_pressure,
            "Data request cannot contain both surface pressure and log of surface pressure parameters.",
        )


filter_registry.register("lnsp_to_sp", LnspToSp)
filter_registry.register("sp_to_lnsp", LnspToSp.reversed)
