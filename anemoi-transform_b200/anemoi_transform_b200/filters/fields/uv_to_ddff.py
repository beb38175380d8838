"""`uv_to_ddff` / `ddff_to_uv` — reference `filters/fields/uv_to_ddff.py:22-131`.

Wind components ↔ speed and meteorological direction, all groups in one device pass
(kernel: AT_EPI_UV2DDFF / AT_EPI_DDFF2UV in csrc/epilogue.cuh).
"""

from __future__ import annotations

from typing import Any, Iterator

from ... import _cabi
from ...matching import MatchingFieldsFilter, MatchingSpec
from . import filter_registry
from .pointwise import NO_COL, device_field, run_epilogue


class WindComponents(MatchingFieldsFilter):
    """Convert U and V wind components to wind speed and direction, and back."""

    MATCHING = MatchingSpec(
        select="param",
        forward=("u_component", "v_component"),
        backward=("wind_speed", "wind_direction"),
    )

    def __init__(
        self,
        *,
        u_component: str = "u",
        v_component: str = "v",
        wind_speed: str = "ws",
        wind_direction: str = "wdir",
        convention: str = "meteo",
        radians: bool = False,
    ) -> None:
        self.u_component = u_component
        self.v_component = v_component
        self.wind_speed = wind_speed
        self.wind_direction = wind_direction
        self.convention = convention
        self.radians = radians
        assert not self.radians, "Radians not (yet) supported"
        if convention != "meteo":
            raise NotImplementedError(f"convention={convention!r}: only 'meteo' is implemented on the device")
        super().__init__()

    # one pair at a time (the reference's extension point) ...
    def forward_transform(self, u_component: Any, v_component: Any) -> Iterator[Any]:
        yield from self.forward_batch([dict(u_component=u_component, v_component=v_component)])[0]

    def backward_transform(self, wind_speed: Any, wind_direction: Any) -> Iterator[Any]:
        yield from self.backward_batch([dict(wind_speed=wind_speed, wind_direction=wind_direction)])[0]

    # ... and every pair of the FieldList in one kernel launch
    def forward_batch(self, groups: list[dict[str, Any]]) -> list[list[Any]]:
        inputs = [g[k] for g in groups for k in ("u_component", "v_component")]
        out = run_epilogue(_cabi.EPI_UV2DDFF, inputs, [NO_COL] * len(inputs))
        return [
            [
                device_field(out, 2 * i, g["u_component"], param=self.wind_speed),
                device_field(out, 2 * i + 1, g["v_component"], param=self.wind_direction),
            ]
            for i, g in enumerate(groups)
        ]

    def backward_batch(self, groups: list[dict[str, Any]]) -> list[list[Any]]:
        inputs = [g[k] for g in groups for k in ("wind_speed", "wind_direction")]
        out = run_epilogue(_cabi.EPI_DDFF2UV, inputs, [NO_COL] * len(inputs))
        return [
            [
                device_field(out, 2 * i, g["wind_speed"], param=self.u_component),
                device_field(out, 2 * i + 1, g["wind_direction"], param=self.v_component),
            ]
            for i, g in enumerate(groups)
        ]


filter_registry.register("uv_to_ddff", WindComponents)
filter_registry.register("ddff_to_uv", WindComponents.reversed)
