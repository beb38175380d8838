"""`cos_sin_from_rad` — reference `filters/fields/cos_sin_from_rad.py:22-126`.

A variable in radians ↔ its cosine and sine (AT_EPI_COSSIN / AT_EPI_ATAN2).  The range check
of the reference (`data.min() < -2π` / `data.max() > 2π` → ValueError, 74-77) runs on the
device (`at_range_flags`); the offending value is fetched only to word the error.
"""

from __future__ import annotations

from typing import Any, Iterator

import numpy as np

from ... import _cabi
from ...batching import fields_to_batch
from ...device import range_flags
from ...matching import MatchingFieldsFilter, MatchingSpec
from . import filter_registry
from ._requests import replace_products
from .pointwise import NO_COL, device_field, run_epilogue


@filter_registry.register("cos_sin_from_rad")
class CosSinFromRad(MatchingFieldsFilter):
    """A filter to convert any variable in radians to cos() and sin() and back."""

    MATCHING = MatchingSpec(select="param", forward=("param",), backward=("cos_param", "sin_param"))

    def __init__(self, *, param: str, cos_param: str | None = None, sin_param: str | None = None) -> None:
        self.param = param
        self.cos_param = cos_param if cos_param is not None else f"cos_{param}"
        self.sin_param = sin_param if sin_param is not None else f"sin_{param}"
        super().__init__()

    def forward_transform(self, param: Any) -> Iterator[Any]:
        yield from self.forward_batch([dict(param=param)])[0]

    def backward_transform(self, cos_param: Any, sin_param: Any) -> Iterator[Any]:
        yield from self.backward_batch([dict(cos_param=cos_param, sin_param=sin_param)])[0]

    def forward_batch(self, groups: list[dict[str, Any]]) -> list[list[Any]]:
        inputs = [g["param"] for g in groups]
        batch = fields_to_batch(inputs)
        flags = range_flags(batch.data, 0, len(inputs), -2 * np.pi, 2 * np.pi)
        for i, f in enumerate(flags):
            if f & 4:  # a NaN makes numpy's min() / max() NaN and both comparisons False
                continue
            if f & 1:
                min = inputs[i].to_numpy().min()  # noqa: A001 - the reference's message names it `min`
                raise ValueError(f"Param {self.param} is expected in radians in the range [-2pi, pi], but {min=}")
            if f & 2:
                max = inputs[i].to_numpy().max()  # noqa: A001
                raise ValueError(f"Param {self.param} is expected in radians in the range [-2pi, pi], but {max=}")
        out = run_epilogue(_cabi.EPI_COSSIN, inputs, [NO_COL] * (2 * len(inputs)), pa=1.0, batch=batch)
        return [
            [
                device_field(out, 2 * i, g["param"], param=self.cos_param),
                device_field(out, 2 * i + 1, g["param"], param=self.sin_param),
            ]
            for i, g in enumerate(groups)
        ]

    def backward_batch(self, groups: list[dict[str, Any]]) -> list[list[Any]]:
        inputs = [g[k] for g in groups for k in ("cos_param", "sin_param")]
        out = run_epilogue(_cabi.EPI_ATAN2, inputs, [NO_COL] * len(groups), pa=1.0, pb=0.0)
        return [[device_field(out, i, g["cos_param"], param=self.param)] for i, g in enumerate(groups)]

    def patch_data_request(self, data_request: dict[str, Any]) -> dict[str, Any]:
        return replace_products(data_request, (self.cos_param, self.sin_param), self.param)
