"""`r_to_d` / `d_to_r` — reference `filters/fields/dewpoint.py:21-75`.

Dewpoint temperature from relative humidity and temperature, and back (AT_EPI_RT2D /
AT_EPI_DT2R: earthkit-meteo's water-phase Tetens formulas, pinned by the reference's golden
vectors tests/field_filters/test_dewpoint.py:23-27).  r == 0 is replaced by 1e-4 inside the
kernel; the input field itself is left untouched.
"""

from __future__ import annotations

from typing import Any, Iterator, Literal

from ... import _cabi
from ...matching import MatchingFieldsFilter, MatchingSpec
from . import filter_registry
from .pointwise import NO_COL, device_field, run_epilogue

EPS = 1.0e-4


class DewPoint(MatchingFieldsFilter):
    """A filter to extract dewpoint temperature from relative humidity and temperature"""

    MATCHING = MatchingSpec(
        select="param",
        forward=("relative_humidity", "temperature"),
        backward=("dewpoint", "temperature"),
    )

    def __init__(
        self,
        *,
        relative_humidity: str = "r",
        temperature: str = "t",
        dewpoint: str = "d",
        return_inputs: Literal["all", "none"] | list[str] = "all",
    ):
        self.return_inputs = return_inputs
        self.relative_humidity = relative_humidity
        self.temperature = temperature
        self.dewpoint = dewpoint
        super().__init__()

    def forward_transform(self, relative_humidity: Any, temperature: Any) -> Iterator[Any]:
        yield from self.forward_batch([dict(relative_humidity=relative_humidity, temperature=temperature)])[0]

    def backward_transform(self, dewpoint: Any, temperature: Any) -> Iterator[Any]:
        yield from self.backward_batch([dict(dewpoint=dewpoint, temperature=temperature)])[0]

    def forward_batch(self, groups: list[dict[str, Any]]) -> list[list[Any]]:
        inputs = [g[k] for g in groups for k in ("relative_humidity", "temperature")]
        out = run_epilogue(_cabi.EPI_RT2D, inputs, [NO_COL] * len(groups))
        return [[device_field(out, i, g["relative_humidity"], param=self.dewpoint)] for i, g in enumerate(groups)]

    def backward_batch(self, groups: list[dict[str, Any]]) -> list[list[Any]]:
        inputs = [g[k] for g in groups for k in ("dewpoint", "temperature")]
        out = run_epilogue(_cabi.EPI_DT2R, inputs, [NO_COL] * len(groups))
        return [[device_field(out, i, g["temperature"], param=self.relative_humidity)] for i, g in enumerate(groups)]


filter_registry.register("r_to_d", DewPoint)
filter_registry.register("d_to_r", DewPoint.reversed)
