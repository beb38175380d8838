"""`q_to_r` / `r_to_q` — reference `filters/fields/q_to_r.py:22-85`.

Specific ↔ relative humidity on pressure levels (pressure = 100 · levelist of the humidity
field forward, of the temperature field backward: q_to_r.py:71,77), all groups in one
device pass (kernel: AT_EPI_QT2R / AT_EPI_RT2Q in csrc/epilogue.cuh).
"""

from __future__ import annotations

from typing import Any, Iterator, Literal

from ... import _cabi
from ...matching import MatchingFieldsFilter, MatchingSpec
from . import filter_registry
from .pointwise import device_field, run_epilogue


class HumidityConversion(MatchingFieldsFilter):
    """Convert specific humidity to relative humidity with standard thermodynamical formulas, and back."""

    MATCHING = MatchingSpec(
        select="param",
        forward=("humidity", "temperature"),
        backward=("relative_humidity", "temperature"),
    )

    def __init__(
        self,
        *,
        relative_humidity: str = "r",
        temperature: str = "t",
        humidity: str = "q",
        return_inputs: Literal["all", "none"] | list[str] = "all",
    ):
        self.return_inputs = return_inputs
        self.relative_humidity = relative_humidity
        self.temperature = temperature
        self.humidity = humidity
        super().__init__()

    def forward_transform(self, humidity: Any, temperature: Any) -> Iterator[Any]:
        yield from self.forward_batch([dict(humidity=humidity, temperature=temperature)])[0]

    def backward_transform(self, relative_humidity: Any, temperature: Any) -> Iterator[Any]:
        yield from self.backward_batch([dict(relative_humidity=relative_humidity, temperature=temperature)])[0]

    def forward_batch(self, groups: list[dict[str, Any]]) -> list[list[Any]]:
        inputs = [g[k] for g in groups for k in ("humidity", "temperature")]
        cols = [(0.0, 0.0, 100 * float(g["humidity"].metadata("levelist")), 0) for g in groups]
        out = run_epilogue(_cabi.EPI_QT2R, inputs, cols)
        return [[device_field(out, i, g["humidity"], param=self.relative_humidity)] for i, g in enumerate(groups)]

    def backward_batch(self, groups: list[dict[str, Any]]) -> list[list[Any]]:
        inputs = [g[k] for g in groups for k in ("relative_humidity", "temperature")]
        # levels are measured in hectopascals
        cols = [(0.0, 0.0, 100 * float(g["temperature"].metadata("levelist")), 0) for g in groups]
        out = run_epilogue(_cabi.EPI_RT2Q, inputs, cols)
        return [[device_field(out, i, g["relative_humidity"], param=self.humidity)] for i, g in enumerate(groups)]


filter_registry.register("q_to_r", HumidityConversion)
filter_registry.register("r_to_q", HumidityConversion.reversed)
