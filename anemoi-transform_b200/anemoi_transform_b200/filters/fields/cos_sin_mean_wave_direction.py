"""`cos_sin_mean_wave_direction` — reference `filters/fields/cos_sin_mean_wave_direction.py:22-128`.

Mean wave direction in degrees ↔ cosine and sine (AT_EPI_COSSIN with the deg2rad factor,
AT_EPI_ATAN2 with the rad2deg factor and the wrap into [0, 360) of lines 96-98).
"""

from __future__ import annotations

from typing import Any, Iterator

import numpy as np

from ... import _cabi
from ...batching import fields_to_batch
from ...matching import MatchingFieldsFilter, MatchingSpec
from . import filter_registry
from ._requests import replace_products
from .pointwise import NO_COL, device_field, run_epilogue


def _unit_factor(ufunc, batch) -> float:
    """numpy's deg2rad / rad2deg multiply by a constant evaluated in the array's own precision;
    ufunc(1) in that precision is that constant."""
    one = np.float32(1.0) if batch.data.element_size() == 4 else np.float64(1.0)
    return float(ufunc(one))


@filter_registry.register("cos_sin_mean_wave_direction")
class CosSinWaveDirection(MatchingFieldsFilter):
    """A filter to convert mean wave direction to cos() and sin() and back."""

    MATCHING = MatchingSpec(
        select="param",
        forward=("mean_wave_direction",),
        backward=("cos_mean_wave_direction", "sin_mean_wave_direction"),
    )

    def __init__(self, *, mean_wave_direction: str = "mwd", cos_mean_wave_direction: str = "cos_mwd", sin_mean_wave_direction: str = "sin_mwd") -> None:
        self.mean_wave_direction = mean_wave_direction
        self.cos_mean_wave_direction = cos_mean_wave_direction
        self.sin_mean_wave_direction = sin_mean_wave_direction
        super().__init__()

    def forward_transform(self, mean_wave_direction: Any) -> Iterator[Any]:
        yield from self.forward_batch([dict(mean_wave_direction=mean_wave_direction)])[0]

    def backward_transform(self, cos_mean_wave_direction: Any, sin_mean_wave_direction: Any) -> Iterator[Any]:
        yield from self.backward_batch([dict(cos_mean_wave_direction=cos_mean_wave_direction, sin_mean_wave_direction=sin_mean_wave_direction)])[0]

    def forward_batch(self, groups: list[dict[str, Any]]) -> list[list[Any]]:
        inputs = [g["mean_wave_direction"] for g in groups]
        batch = fields_to_batch(inputs)
        out = run_epilogue(_cabi.EPI_COSSIN, inputs, [NO_COL] * (2 * len(inputs)), pa=_unit_factor(np.deg2rad, batch), batch=batch)
        return [
            [
                device_field(out, 2 * i, g["mean_wave_direction"], param=self.cos_mean_wave_direction),
                device_field(out, 2 * i + 1, g["mean_wave_direction"], param=self.sin_mean_wave_direction),
            ]
            for i, g in enumerate(groups)
        ]

    def backward_batch(self, groups: list[dict[str, Any]]) -> list[list[Any]]:
        inputs = [g[k] for g in groups for k in ("cos_mean_wave_direction", "sin_mean_wave_direction")]
        batch = fields_to_batch(inputs)
        out = run_epilogue(_cabi.EPI_ATAN2, inputs, [NO_COL] * len(groups), pa=_unit_factor(np.rad2deg, batch), pb=1.0, batch=batch)
        return [[device_field(out, i, g["cos_mean_wave_direction"], param=self.mean_wave_direction)] for i, g in enumerate(groups)]

    def patch_data_request(self, data_request: dict[str, Any]) -> dict[str, Any]:
        return replace_products(data_request, (self.cos_mean_wave_direction, self.sin_mean_wave_direction), self.mean_wave_direction)
