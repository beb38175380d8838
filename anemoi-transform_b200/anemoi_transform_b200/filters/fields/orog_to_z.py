"""`orog_to_z_fields` / `z_to_orog_fields` — reference `filters/fields/orog_to_z.py:19-98`.

Orography (m) ↔ surface geopotential (m²/s²): `orog · g` forward, `z / g` backward
(AT_EPI_AFFINE / AT_EPI_AFFINE_INV with a zero offset; bit-exact except that a −0.0 product
comes out as +0.0).
"""

from __future__ import annotations

from typing import Any

from ... import _cabi
from ...constants import g_gravitational_acceleration
from ...filter import SingleFieldFilter
from . import filter_registry
from ._requests import rename_on_levels
from .pointwise import NO_COL, device_field, run_epilogue


class Orography(SingleFieldFilter):
    r"""A filter to convert orography in m to surface geopotential in m²/s², and back."""

    optional_inputs = {"orography": "orog", "geopotential": "z"}

    def forward_select(self):
        return {"param": self.orography}

    def backward_select(self):
        return {"param": self.geopotential}

    def forward_transform(self, orography: Any) -> Any:
        return self.forward_transform_batch([orography])[0]

    def backward_transform(self, geopotential: Any) -> Any:
        return self.backward_transform_batch([geopotential])[0]

    def forward_transform_batch(self, fields: list[Any]) -> list[Any]:
        out = run_epilogue(_cabi.EPI_AFFINE, fields, [NO_COL] * len(fields), pa=g_gravitational_acceleration, pb=0.0)
        return [device_field(out, i, f, param=self.geopotential) for i, f in enumerate(fields)]

    def backward_transform_batch(self, fields: list[Any]) -> list[Any]:
        out = run_epilogue(_cabi.EPI_AFFINE_INV, fields, [NO_COL] * len(fields), pa=g_gravitational_acceleration, pb=0.0)
        return [device_field(out, i, f, param=self.orography) for i, f in enumerate(fields)]

    def patch_data_request(self, data_request: Any) -> Any:
        return rename_on_levels(data_request, self.geopotential, self.orography, "Data request cannot contain both orography and geopotential parameters.")


filter_registry.register("orog_to_z_fields", Orography)
filter_registry.register("z_to_orog_fields", Orography.reversed)
